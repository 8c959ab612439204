"""
-m gpu: the tensor-core (tcgen05, bf16 operands / fp32 accumulate) mode against the fp32 CPU oracle.
Stated tolerance (include/sfgpi.h, SURVEY section 7 "fp32 1e-5 parity vs tensor cores"): BF16_TOL = 2e-2 scale-relative
on psi / q, relative Frobenius error < 1e-2; GPI argmax must agree wherever the fp32 top-1/top-2 gap exceeds
BF16_TOL * max|q| (ties inside tolerance may flip).
"""
import numpy as np
import pytest
import torch

from oracle.sf_oracle import OracleSF, synthetic_transitions
from tests.golden_util import rel_err
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2
HYPER_BF16 = dict(gu.HYPER, precision='bf16')


def fro_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make(S, A, D, N, seed, tsf_dim=None, beta=1):
    gen = torch.Generator().manual_seed(seed)
    o = OracleSF(S, A, D, (256, 256), ('relu', 'relu'), tsf_dim=tsf_dim, beta=beta)
    for _ in range(N):
        o.add_random_policy(gen)
    return o, gen


@pytest.mark.parametrize('S,A,D,N,B,hopper', [
    (4, 9, 12, 4, 4096, False),          # 128 tiles: one tile per CTA (no ping-pong partner)
    (4, 9, 12, 6, 33 * 128 - 5, False),  # 198 tiles: paired mode, odd tile count per policy, ragged last tile
    (11, 27, 50, 3, 1000, True),         # Hopper: output layer in 6 chunks of <= 256 columns, S = 11
    (4, 2, 20, 3, 32, False),            # CartPole, a single partial tile
])
def test_bf16_forward_gpi_vs_oracle(S, A, D, N, B, hopper):
    meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N)
    o, gen = make(S, A, D, N, seed=21)
    sf = gu.build_g2(meta, oracle=o, hyper=HYPER_BF16)
    assert sf._library.precision == 'bf16'
    x = synthetic_transitions(B, S, A, D, gen, hopper=hopper)[0]
    psi_ref = o.get_successors(x)
    psi = sf.get_successors(x.cuda()).cpu()
    assert rel_err(psi, psi_ref) < BF16_TOL and fro_err(psi, psi_ref) < 1e-2
    psi_t = sf._library.forward_psi(x.cuda(), 1, 1, target=True).cpu()
    assert fro_err(psi_t[:, 0], psi_ref[:, 1]) < 1e-2
    q_ref, task_ref = o.GPI(x, 1)
    q, task = sf.GPI(x.cuda(), 1)
    assert rel_err(q.cpu(), q_ref) < BF16_TOL
    ok, nbad = gu.argmax_mismatch_ok(q_ref, task_ref, task.cpu(), 'task', BF16_TOL)
    assert ok, f'{nbad} task mismatches outside tolerance'
    _, key_a, key_t = sf._library.gpi(x.cuda(), sf.fit_w[1].weight, want_q=False)
    act, val = sf._library.decode_keys(key_a, want_value=True)
    act_ref = torch.argmax(torch.max(q_ref, dim=1).values, dim=-1)
    ok, nbad = gu.argmax_mismatch_ok(q_ref, act_ref, act.cpu(), 'action', BF16_TOL)
    assert ok, f'{nbad} action mismatches outside tolerance'
    # self-consistency (exact): keys agree with the q the same kernel returned
    assert torch.equal(val, q.reshape(B, -1).max(dim=1).values)
    assert torch.equal(act, torch.argmax(q.max(dim=1).values, dim=-1))
    assert torch.equal(sf._library.decode_keys(key_t), torch.argmax(q.max(dim=2).values, dim=1))


@pytest.mark.parametrize('variant,N', [('g2', 4), ('g3', 6)])
def test_bf16_train_step_vs_oracle(variant, N):
    S, A, D, B = 4, 9, 12, 4096
    tsf = variant == 'g3'
    meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N, gdim=100, beta=1, use_gpi=True)
    o, gen = make(S, A, D, N, seed=31, tsf_dim=100 if tsf else None)
    if tsf:
        sf, ag = gu.build_g3(meta, oracle=o)
        sf._library.set_precision('bf16')
    else:
        sf = ag = gu.build_g2(meta, oracle=o, hyper=HYPER_BF16)
    W_before = [W.clone() for W, _ in gu.psi_params(sf, 1)]
    tr = synthetic_transitions(B, S, A, D, gen)
    ref = o.tsf_update_successor(tr, 1, True) if tsf else o.update_successor(tr, 1, True)
    out = ag.update_successor(gu.cuda_tr(tr), 1, True)
    assert np.allclose([float(v) for v in out], [float(v) for v in ref], rtol=3e-2, atol=1e-6)
    for l, (W, _) in enumerate(gu.psi_params(sf, 1)):
        d_mine, d_ref = (W - W_before[l]).flatten().double(), (o.psi[1][l][0] - W_before[l]).flatten().double()
        cos = float(d_mine @ d_ref / (d_mine.norm() * d_ref.norm()))
        assert cos > 0.9, f'layer {l}: update direction cos {cos}'
    # all-task step with GPI over all reward vectors (n_w = N: the multi-vector epilogue)
    tr = synthetic_transitions(B, S, A, D, gen)
    ref = o.ensemble_update_frozen(tr, tsf=tsf, use_gpi=True)
    losses = ag.update_successor_all(gu.cuda_tr(tr), use_gpi=True).cpu()
    for i in range(N):
        assert np.allclose(losses[i].numpy(), [float(v) for v in ref[i]], rtol=3e-2, atol=1e-6)
