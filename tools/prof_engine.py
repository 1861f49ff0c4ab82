"""Driver for ncu: one eager forward of the fused engine on BERT-base (random init, one calibration batch)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'transformer-quantization_b200'))
import bench
from engine.fused import FusedBertEngine
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
ids = bench.synthetic_ids(1234)[0].to(dev)
mask = torch.ones_like(ids)
with torch.no_grad():
    model(ids, mask); model.fix_ranges(); model(ids, mask)
eng = FusedBertEngine(model, bench.BATCH, bench.SEQ)
for _ in range(3):
    eng(ids, mask)
torch.cuda.synchronize()
# ncu --profile-from-start off: only this forward is captured (5 kernels per encoder layer)
torch.cuda.profiler.start()
eng(ids, mask)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
