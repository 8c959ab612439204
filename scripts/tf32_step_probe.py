"""Where does the tf32x3 train step differ from the golden / oracle step (developer probe)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.golden_util import load, transitions, rel_err, t
from tests import gpu_util as gu

for precision in ('fp32', 'tf32x3'):
    meta, z = load('g2_reacher_h256')
    sf = gu.build_g2(meta, z, hyper=dict(gu.HYPER, precision=precision))
    i = meta['policy']
    for k in range(meta['K']):
        out = sf.update_successor(gu.cuda_tr(transitions(z, k)), i, meta['use_gpi'])
        print(precision, 'losses', [float(v) for v in out], z['out.losses'][k])
    lib = sf._library
    ms, vs = lib.spec.views(lib.m[i].cpu()), lib.spec.views(lib.v[i].cpu())
    for l, (W, b) in enumerate(gu.psi_params(sf, i)):
        mref = t(z[f'post.adam.W{l}.m'])
        dm = (ms[l][0] - mref).abs()
        dW = (W - t(z[f'post.psi.W{l}'])).abs()
        j = int(dW.argmax())
        print(f'{precision} L{l}: W rel_err {rel_err(W, z[f"post.psi.W{l}"]):.2e} at flat {j}: m_mine {float(ms[l][0].flatten()[j]):.3e} m_ref {float(mref.flatten()[j]):.3e}'
              f' | m rel_err {rel_err(ms[l][0], mref):.2e} max|m| {float(mref.abs().max()):.2e} max abs dm {float(dm.max()):.2e}'
              f' | b rel_err {rel_err(b, z[f"post.psi.b{l}"]):.2e}')
