#!/usr/bin/env python
"""BASELINE-size golden: the UNMODIFIED reference (``models/quantized_bert.py`` + ``quantization/*`` through
baseline/reference_arm.py) on the host CPU, on exactly bench.py's model -- BERT-base, 12 layers, B=32, T=128,
weights seed 0 (oracle.bert_oracle.random_bert_state_dict), token ids seed 1234, W8 symmetric / A8 asymmetric,
ONE calibration batch (the same batch), fix_ranges, eval forward.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_fullsize.py        (build container, ~1 min)

Writes tests/golden/bert_base_fullsize.npz: the logits, and for each of the 161 activation-quantizer sites (in
the reference's module order) its (delta, zero_float) and a strided SAMPLE of its output (every 2053rd element
of the flattened tensor, 2053 prime: the samples walk through all rows, heads and columns) -- 689 M elements
would not fit a fixture, 336 k samples do.  tests/test_gpu_fullsize_parity.py compares the module path on the
B200 against it, site by site.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
STRIDE = 2053


def main():
    from baseline import reference_arm as RA
    from oracle.bert_oracle import random_bert_state_dict
    assert RA.available(), 'install the reference first: tools/install_reference.sh'
    torch.set_num_threads(os.cpu_count() or 1)
    sd = random_bert_state_dict(seed=0)
    ids = torch.randint(0, 30522, (32, 128), generator=torch.Generator().manual_seed(1234))
    mask = torch.ones_like(ids)
    ref = RA.ReferenceBert(sd)
    sites = [(n, m) for n, m in ref.model.named_modules()
             if n.endswith('activation_quantizer') and hasattr(m, 'quantizer')]      # (FP32Acts: embedding tables)
    out = {'ids': ids.numpy(), 'stride': np.array(STRIDE), 'n_sites': np.array(len(sites))}
    with torch.no_grad():
        ref(ids, mask)
        ref.fix_ranges()
        got = {}
        hooks = [m.register_forward_hook(lambda mod, inp, o, i=i: got.__setitem__(i, o.detach().reshape(-1)[::STRIDE].clone()))
                 for i, (n, m) in enumerate(sites)]
        logits = ref(ids, mask)
        for h in hooks:
            h.remove()
    out['logits'] = logits.numpy()
    for i, (n, m) in enumerate(sites):
        q = m.quantizer
        out[f'q{i}.name'] = np.array(n)
        out[f'q{i}.delta'] = q._delta.detach().numpy().reshape(-1)
        out[f'q{i}.zero_float'] = q._zero_float.detach().numpy().reshape(-1)
        out[f'q{i}.sample'] = got[i].numpy()
    np.savez_compressed(os.path.join(HERE, 'bert_base_fullsize.npz'), **out)
    print('wrote bert_base_fullsize.npz:', len(sites), 'sites, logits', logits.min().item(), logits.max().item())


if __name__ == '__main__':
    main()
