"""Determinism stress of the encoder chain kernel: the BERT-base engine forward N times on fresh random ids, every buffer
compared with the per-stage kernels (TQ_ENGINE_CHAIN=0) -- any stale read across a stage boundary shows as a mismatch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import bench
from engine.fused import FusedBertEngine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
ids = bench.synthetic_ids(1234)[0].to(dev)
mask = torch.ones_like(ids)
with torch.no_grad():
    model(ids, mask); model.fix_ranges()
os.environ['TQ_ENGINE_CHAIN'] = '0'
ref = FusedBertEngine(model, bench.BATCH, bench.SEQ)
os.environ['TQ_ENGINE_CHAIN'] = '1'
eng = FusedBertEngine(model, bench.BATCH, bench.SEQ)
assert eng.chain == 1 and ref.chain == 0
g = torch.Generator(device='cpu').manual_seed(7)
bad = 0
for it in range(N):
    x = torch.randint(0, 30000, (bench.BATCH, bench.SEQ), generator=g).to(dev)
    a = ref(x, mask).clone(); ra = (ref.x8.clone(), ref.a8.clone(), ref.f8.clone(), ref.qkv.clone())
    b = eng(x, mask).clone(); rb = (eng.x8, eng.a8, eng.f8, eng.qkv)
    torch.cuda.synchronize()
    if not (torch.equal(a, b) and all(torch.equal(p, q) for p, q in zip(ra, rb))):
        bad += 1
print(f'stress_chain: {N} forwards, {bad} mismatching')
sys.exit(1 if bad else 0)
