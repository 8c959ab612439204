"""
-m gpu: the tensor-core modes at the reference's precision (tcgen05 kind::tf32 on fp32 operands, csrc/mlp_stream_tc.cu).
  'tf32x3'  3-pass split, fp32 accumulation: held to the SAME bounds as the fp32 CUDA-core mode -- FWD_TOL 1e-5 scale-relative
            on psi / q / losses, STEP_TOL 1e-4 on post-step weights, GPI argmax exact except ties inside the tolerance -- against
            the committed outputs of the unmodified reference (tests/golden) and the CPU oracle.
  'tf32'    one pass: stated tolerance TF32_TOL = 2e-3 scale-relative on psi / q (3.7e-4 .. 4.7e-4 measured in
            CPU emulation), argmax equal wherever the fp32 top-1 / top-2 gap exceeds it.
"""
import numpy as np
import pytest
import torch

from oracle.sf_oracle import OracleSF, synthetic_transitions
from tests.golden_util import load, oracle_from_golden, transitions, n_layers, rel_err, t
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu
FWD_TOL, STEP_TOL, TF32_TOL = 1e-5, 1e-4, 2e-3
TOL = {'tf32x3': FWD_TOL, 'tf32': TF32_TOL}


def make_oracle(S, A, D, N, seed, tsf_dim=None, beta=1):
    gen = torch.Generator().manual_seed(seed)
    o = OracleSF(S, A, D, (256, 256), ('relu', 'relu'), tsf_dim=tsf_dim, beta=beta)
    for _ in range(N):
        o.add_random_policy(gen)
    return o, gen


@pytest.mark.parametrize('precision', ['tf32x3', 'tf32'])
def test_forward_gpi_vs_golden_h256(precision):
    """The reference's own outputs (g2_reacher_h256: full-width MLP): psi of every policy, q, task, batch-1 squeeze."""
    meta, z = load('g2_reacher_h256')
    sf = gu.build_g2(meta, z, hyper=dict(gu.HYPER, precision=precision))
    tol = TOL[precision]
    x = t(z['tr0.states']).cuda()
    assert rel_err(sf.get_successors(x).cpu(), z['out.psi_all']) < tol
    q, task = sf.GPI(x, meta['policy'])
    assert rel_err(q.cpu(), z['out.q']) < tol
    ok, nbad = gu.argmax_mismatch_ok(t(z['out.q']), t(z['out.task']), task.cpu(), 'task', tol)
    assert ok and (precision == 'tf32' or nbad == 0)
    q1, task1 = sf.GPI(x[:1], meta['policy'])
    assert tuple(q1.shape) == tuple(z['out.q_b1'].shape) and task1.dim() == 0


@pytest.mark.parametrize('precision', ['tf32x3', 'tf32'])
@pytest.mark.parametrize('S,A,D,N,B,hopper', [
    (4, 9, 12, 4, 4096, False),          # Reacher, the bench's forward sizes
    (4, 9, 12, 3, 5 * 128 - 7, False),   # ragged last tile
    (11, 27, 50, 3, 1000, True),         # Hopper: psi output in 6 chunks (A re-produced per chunk), S = 11 -> two K = 8 steps
    (4, 2, 20, 3, 32, False),            # CartPole, a single partial tile
    (4, 9, 12, 40, 300, False),          # more units than SMs: the persistent loop takes several units per CTA
])
def test_forward_gpi_vs_oracle(precision, S, A, D, N, B, hopper):
    tol = TOL[precision]
    meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N)
    o, gen = make_oracle(S, A, D, N, seed=21)
    sf = gu.build_g2(meta, oracle=o, hyper=dict(gu.HYPER, precision=precision))
    assert sf._library.precision == precision
    x = synthetic_transitions(B, S, A, D, gen, hopper=hopper)[0]
    psi_ref = o.get_successors(x)
    assert rel_err(sf.get_successors(x.cuda()).cpu(), psi_ref) < tol
    assert rel_err(sf._library.forward_psi(x.cuda(), 1, 1, target=True).cpu()[:, 0], psi_ref[:, 1]) < tol
    q_ref, task_ref = o.GPI(x, 1)
    q, task = sf.GPI(x.cuda(), 1)
    assert rel_err(q.cpu(), q_ref) < tol
    ok, nbad = gu.argmax_mismatch_ok(q_ref, task_ref, task.cpu(), 'task', tol)
    assert ok, f'{nbad} task mismatches outside tolerance'
    _, key_a, key_t = sf._library.gpi(x.cuda(), sf.fit_w[1].weight, want_q=False)
    act, val = sf._library.decode_keys(key_a, want_value=True)
    act_ref = torch.argmax(torch.max(q_ref, dim=1).values, dim=-1)
    ok, nbad = gu.argmax_mismatch_ok(q_ref, act_ref, act.cpu(), 'action', tol)
    assert ok and (precision == 'tf32' or nbad <= max(1, B // 1000)), f'{nbad} action mismatches'
    assert rel_err(val.cpu(), q_ref.reshape(B, -1).max(dim=1).values) < tol
    # self-consistency (exact): keys agree with the q the same kernel returned
    assert torch.equal(val, q.reshape(B, -1).max(dim=1).values)
    assert torch.equal(act, torch.argmax(q.max(dim=1).values, dim=-1))


@pytest.mark.parametrize('nw', [5, 13, 40])
def test_tf32x3_multi_vector_gpi_vs_oracle(nw):
    """Folded GPI under n_w reward vectors (blocked column order, several 256-column chunks) against the oracle's psi . w."""
    import ctypes as C
    from deep_successor_features_for_transfer_b200 import _lib
    from deep_successor_features_for_transfer_b200.library import _stream
    S, A, D, N, B = 4, 9, 12, 3, 500
    meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N)
    o, gen = make_oracle(S, A, D, N, seed=47)
    sf = gu.build_g2(meta, oracle=o, hyper=dict(gu.HYPER, precision='tf32x3'))
    lib = sf._library
    x = synthetic_transitions(B, S, A, D, gen)[0]
    w = (torch.rand(nw, D, generator=gen) * 0.02 - 0.01).contiguous()
    lib._pack('online', 0, N)
    ka = torch.empty(nw, B, dtype=torch.int64, device='cuda')
    kt = torch.empty(nw, B, dtype=torch.int64, device='cuda')
    _lib.call('sfgpi_keys_fill', ka.data_ptr(), ka.numel(), _stream())
    _lib.call('sfgpi_keys_fill', kt.data_ptr(), kt.numel(), _stream())
    xd, wd = x.cuda(), w.cuda()
    a = lib._fwd_args(lib.online, 0, N, xd)
    a.w, a.n_w, a.w_diag, a.task_base = wd.data_ptr(), nw, 0, 0
    a.key_action, a.key_task = ka.data_ptr(), kt.data_ptr()
    lib._forward(a, 'online', fresh=True)
    torch.cuda.synchronize()
    act, val = lib.decode_keys(ka, want_value=True)
    task = lib.decode_keys(kt)
    q_all = torch.einsum('bnad,wd->wbna', o.get_successors(x), w)
    scale = float(q_all.abs().max())
    for wi in range(nw):
        q = q_all[wi]
        assert float((val[wi].cpu() - q.reshape(B, -1).max(dim=1).values).abs().max()) < FWD_TOL * scale
        ok, nb = gu.argmax_mismatch_ok(q, torch.argmax(q.max(dim=1).values, dim=-1), act[wi].cpu(), 'action', FWD_TOL)
        assert ok and nb <= 1, f'vector {wi}: {nb} action mismatches'
        ok, nb = gu.argmax_mismatch_ok(q, torch.argmax(q.max(dim=2).values, dim=1), task[wi].cpu(), 'task', FWD_TOL)
        assert ok and nb <= 1, f'vector {wi}: {nb} task mismatches'
