"""Post-step weight error statistics of the fp32 (FFMA) and tf32x3 modes against the fp32 oracle (developer probe)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests import test_gpu_tf32 as tt
from tests import gpu_util as gu
from oracle.sf_oracle import synthetic_transitions

def stats(W, R):
    W, R = W.double(), R.double()
    d = (W - R).abs() / R.abs().max()
    return f'mean {float(d.mean()):.1e} max {float(d.max()):.1e} frac>1e-4 {float((d > 1e-4).double().mean()):.1e} frac>1e-3 {float((d > 1e-3).double().mean()):.1e}'

TSF = len(sys.argv) > 1 and sys.argv[1] == 'tsf'
for precision in ('fp32', 'tf32x3'):
    for (B, N, seed, pols) in [(4096, 4, 6, [0, 3, 0]), (1000, 5, 9, None)]:
        meta = dict(S=4, A=9, D=12, hidden=[256, 256], acts=['relu', 'relu'], N=N, gdim=100, beta=1 if pols else 30, use_gpi=True)
        o, gen = tt.make_oracle(4, 9, 12, N, seed=seed, tsf_dim=100 if TSF else None, beta=meta['beta'])
        if TSF:
            sf, ag = gu.build_g3(meta, oracle=o)
            sf._library.set_precision(precision)
        else:
            sf = ag = gu.build_g2(meta, oracle=o, hyper=dict(gu.HYPER, precision=precision))
        if pols is not None:
            for pol in pols:
                tr = synthetic_transitions(B, 4, 9, 12, gen)
                ref = o.tsf_update_successor(tr, pol, True) if TSF else o.update_successor(tr, pol, True)
                out = ag.update_successor(gu.cuda_tr(tr), pol, True)
                print(precision, 'losses', [float(v) for v in out], [float(v) for v in ref])
            check = [0, 3]
        else:
            for k in range(2):
                tr = synthetic_transitions(B, 4, 9, 12, gen)
                o.ensemble_update_frozen(tr, tsf=TSF, use_gpi=True)
                ag.update_successor_all(gu.cuda_tr(tr), use_gpi=True)
            check = list(range(N))
        lib = sf._library
        for pol in check:
            for l, (W, b) in enumerate(gu.psi_params(sf, pol)):
                m = lib.spec.views(lib.m[pol].cpu())[l][0]
                print(precision, f'B{B} pol{pol} L{l}:', stats(W, o.psi[pol][l][0]), '| m fro', f'{tt.fro_err(m, o.adam[pol]["m"]["sf"][2 * l]):.1e}', flush=True)
