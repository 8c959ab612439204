"""Prints per-layer relative Frobenius errors of PackedSFLibrary.psi_gradients (both precision modes) vs torch autograd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.synthetic import synthetic_transitions
from tests import gpu_util as gu
from tests.test_gpu_bf16 import make, fro_err, torch_psi_grads

for (S, A, D, N, B, hopper) in [(4, 9, 12, 3, 1000, False), (4, 9, 12, 5, 33 * 128 - 5, False), (11, 27, 50, 2, 300, True), (4, 2, 20, 2, 32, False)]:
    for precision in ('fp32', 'bf16'):
        meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N)
        o, gen = make(S, A, D, N, seed=77)
        sf = gu.build_g2(meta, oracle=o, hyper=dict(gu.HYPER, precision=precision))
        lib = sf._library
        tr = synthetic_transitions(B, S, A, D, gen, hopper=hopper)
        x, actions = tr[0], tr[1]
        lo, n_pol = 1, N - 1
        d_out = torch.randn(n_pol, B, D, generator=gen) * 1e-4
        ref = torch_psi_grads(o, lo, n_pol, x, actions, d_out, emulate_bf16=precision == 'bf16')
        got = lib.psi_gradients(x.cuda(), actions.cuda(), d_out.cuda(), lo, n_pol).cpu()
        torch.cuda.synchronize()
        for p in range(n_pol):
            errs = []
            for l, ((W, b), (gW, gb)) in enumerate(zip(lib.spec.views(got[p]), ref[p])):
                errs.append(f'L{l}: dW {fro_err(W, gW):.2e} db {fro_err(b, gb):.2e}')
            print(f'S={S} A={A} D={D} B={B} {precision} pol{p}: ' + ' | '.join(errs), flush=True)
