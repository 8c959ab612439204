# -*- coding: UTF-8 -*-
"""
Measures what tests/test_gpu_bf16.py bounds: K-step drift of the tensor-core train step against the emulating oracle and the
fp32 oracle (kstep_metrics), and the 1001-step loss trajectories (kernel vs fp32 oracle, emulating oracle vs fp32 oracle,
kernel vs emulating oracle).  Output: gpurun_out/bf16_drift.json -> summarised in profiles/r02_bf16_drift.md.
    python scripts/bf16_drift_probe.py [precision]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tests import test_gpu_bf16 as tb  # noqa: E402
from tests import gpu_util as gu  # noqa: E402
from oracle.sf_oracle import synthetic_transitions  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
out = {'precision': precision, 'kstep': {}}
for variant, N, K in [('g2', 4, 1), ('g3', 4, 1), ('g3', 4, 10), ('g2', 3, 10)]:
    out['kstep'][f'{variant}_N{N}_K{K}'] = tb.kstep_metrics(variant, N, K, precision=precision)
    print(variant, N, K, out['kstep'][f'{variant}_N{N}_K{K}'], flush=True)

S, A, D, N, B, K = 4, 9, 12, 2, 256, 1001
meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N, gdim=100, beta=1, use_gpi=True, target_update_ev=1000)
o, gen = tb.make(S, A, D, N, seed=61, tsf_dim=100)
oe, _ = tb.make(S, A, D, N, seed=61, tsf_dim=100)
oe.emulate = {'bf16': 'bf16', 'tf32': 'tf32'}.get(precision)
o.target_update_ev = oe.target_update_ev = 1000
sf, ag = gu.build_g3(meta, oracle=o)
sf._library.set_precision(precision)
batches = [synthetic_transitions(B, S, A, D, gen) for _ in range(16)]
dev_b = [gu.cuda_tr(b) for b in batches]
got, ref, emu = [], [], []
for k in range(K):
    got.append(torch.stack(list(ag.update_successor(dev_b[k % 16], 1, True))))
    ref.append([float(x) for x in o.tsf_update_successor(batches[k % 16], 1, True)])
    emu.append([float(x) for x in oe.tsf_update_successor(batches[k % 16], 1, True)])
got = torch.stack(got).cpu().double().numpy()
ref, emu = np.array(ref), np.array(emu)
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-9)
for name, d in (('kernel_vs_fp32', rel(got, ref)), ('emul_vs_fp32', rel(emu, ref)), ('kernel_vs_emul', rel(got, emu))):
    out[name] = {'max': float(d.max()), 'median': float(np.median(d)), 'p99': float(np.quantile(d, 0.99)),
                 'max_first_100': float(d[:100].max()), 'argmax_step': int(d.max(axis=1).argmax())}
    print(name, out[name], flush=True)
out['loss_first_last'] = {'fp32': [ref[0].tolist(), ref[-1].tolist()], 'kernel': [got[0].tolist(), got[-1].tolist()]}
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', f'{precision}_drift.json'), 'w') as f:
    json.dump(out, f, indent=1)
