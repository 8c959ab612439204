"""
-m gpu: the parity tests proper.  Every call goes Python API -> ctypes -> C ABI (libsfgpi.so) -> sm_100a kernels and is
compared with (a) the committed outputs of the unmodified reference (tests/golden) and (b) the CPU oracle on the same
seeded inputs.  Tolerances (fp32 mode, BASELINE north_star: 1e-5 relative):
  FWD_TOL  1e-5  scale-normalised max error on psi / q / losses
  STEP_TOL 1e-4  on post-step weights after <= 4 Adam steps (Adam's m/(sqrt(v)+eps) turns a 1e-7 relative gradient
                 difference on near-zero gradients into a visible fraction of lr; the mean error stays < 1e-6)
  GPI argmax: bit-exact except where the reference's top-1/top-2 gap is below FWD_TOL * max|q| (ties inside tolerance).
"""
import os

import numpy as np
import pytest
import torch

from oracle.sf_oracle import OracleSF, synthetic_transitions
from tests.golden_util import load, oracle_from_golden, transitions, n_layers, rel_err, t
from tests import gpu_util as gu

pytestmark = pytest.mark.gpu
FWD_TOL, STEP_TOL = 1e-5, 1e-4
G2 = ['g2_reacher_gpi', 'g2_reacher_nogpi', 'g2_reacher_sync', 'g2_reacher_h256', 'g2_cartpole_tanh', 'g2_hopper_gpi']


def mean_err(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().mean() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize('name', G2)
def test_g2_forward_gpi_vs_golden(name):
    meta, z = load(name)
    sf = gu.build_g2(meta, z)
    x = t(z['tr0.states']).cuda()
    psi = sf.get_successors(x)
    assert psi.shape == tuple(z['out.psi_all'].shape)
    assert rel_err(psi.cpu(), z['out.psi_all']) < FWD_TOL
    assert rel_err(sf.get_successor(x, 1).cpu(), z['out.psi_all'][:, 1]) < FWD_TOL
    q, task = sf.GPI(x, meta['policy'])
    assert rel_err(q.cpu(), z['out.q']) < FWD_TOL
    ok, nbad = gu.argmax_mismatch_ok(t(z['out.q']), t(z['out.task']), task.cpu(), 'task', FWD_TOL)
    assert ok and nbad <= 1
    q1, task1 = sf.GPI(x[:1], meta['policy'])
    assert q1.shape == tuple(z['out.q_b1'].shape) and task1.dim() == 0 and task1.dtype == torch.int64
    assert int(task1) == int(z['out.task_b1'])
    # unfused epilogue on the materialised psi gives the same answer as the fused one
    q2, _, kt = sf._library.gpi_from_psi(psi, sf.fit_w[meta['policy']].weight)
    assert rel_err(q2.cpu(), z['out.q']) < FWD_TOL
    assert torch.equal(sf._library.decode_keys(kt).cpu(), task.cpu().reshape(-1))


@pytest.mark.parametrize('name', G2)
def test_g2_update_successor_vs_golden(name):
    meta, z = load(name)
    sf = gu.build_g2(meta, z)
    i = meta['policy']
    for k in range(meta['K']):
        out = sf.update_successor(gu.cuda_tr(transitions(z, k)), i, meta['use_gpi'])
        assert isinstance(out, tuple) and len(out) == 3 and out[0].dim() == 0
        assert np.allclose([float(v) for v in out], z['out.losses'][k], rtol=2e-5, atol=1e-8)
    post, tgt = gu.psi_params(sf, i), gu.psi_params(sf, i, target=True)
    st = sf.psi[i][0][2].state
    lins = gu.linears(sf.psi[i][0][0].net)
    for l in range(n_layers(meta)):
        assert rel_err(post[l][0], z[f'post.psi.W{l}']) < STEP_TOL and mean_err(post[l][0], z[f'post.psi.W{l}']) < 2e-6
        assert rel_err(post[l][1], z[f'post.psi.b{l}']) < STEP_TOL
        assert rel_err(tgt[l][0], z[f'post.tgt.W{l}']) < STEP_TOL
        assert rel_err(st[lins[l].weight]['exp_avg'].cpu(), z[f'post.adam.W{l}.m']) < 2e-5
        assert rel_err(st[lins[l].weight]['exp_avg_sq'].cpu(), z[f'post.adam.W{l}.v']) < 4e-5
        assert rel_err(st[lins[l].bias]['exp_avg'].cpu(), z[f'post.adam.b{l}.m']) < 2e-5
    assert rel_err(sf.fit_w[i].weight.data.cpu(), z['post.w']) < STEP_TOL
    assert int(sf._library.step[i]) == int(z['post.adam.w.step'])
    assert sf.updates_since_target_updated == list(z['post.updates_since_target_updated'])
    for j in range(meta['N']):                                   # untouched policies stay bit-identical
        if j != i:
            for l, (W, b) in enumerate(gu.psi_params(sf, j)):
                assert torch.equal(W, t(z[f'init.psi{j}.W{l}']))
    assert sf.update_successor(None, i) is None                 # empty replay -> None (sfdqn.py:304-305)


@pytest.mark.parametrize('name', ['g3_reacher_gpi', 'g3_reacher_beta30'])
def test_g3_tsf_vs_golden(name):
    meta, z = load(name)
    dsf, ag = gu.build_g3(meta, z)
    assert rel_err(dsf.get_next_successors(t(z['tr0.states']).cuda()).cpu(), z['out.next_psi_all']) < FWD_TOL
    for k in range(meta['K']):
        out = ag.update_successor(gu.cuda_tr(transitions(z, k)), meta['policies'][k], meta['use_gpi'])
        assert np.allclose([float(v) for v in out], z['out.losses'][k], rtol=2e-5, atol=1e-8)
    for i in sorted(set(meta['policies'])):
        for l, (W, b) in enumerate(gu.psi_params(dsf, i)):
            assert rel_err(W, z[f'post.psi{i}.W{l}']) < STEP_TOL
            assert rel_err(b, z[f'post.psi{i}.b{l}']) < STEP_TOL
        assert rel_err(dsf.fit_w[i].weight.data.cpu(), z[f'post.w{i}']) < STEP_TOL
        assert rel_err(ag.g_functions[i].weight.data.cpu(), z[f'post.g{i}.W']) < STEP_TOL
        assert rel_err(ag.g_functions[i].bias.data.cpu(), z[f'post.g{i}.b']) < STEP_TOL
        st = dsf.psi[i][0][2].state
        assert rel_err(st[ag.h_function.weight]['exp_avg'].cpu(), z[f'post.adam{i}.hW.m']) < 2e-5
        assert rel_err(st[ag.g_functions[i].weight]['exp_avg_sq'].cpu(), z[f'post.adam{i}.gW.v']) < 4e-5
    assert rel_err(ag.h_function.weight.data.cpu(), z['post.h.W']) < STEP_TOL
    assert rel_err(ag.h_function.bias.data.cpu(), z['post.h.b']) < STEP_TOL
    with pytest.raises(Exception):
        dsf.update_successor(None, 0)                            # 'This function should not be called'


def test_g1_ensemble_vs_golden():
    from deep_successor_features_for_transfer_b200.ensemble import DeepSF as DeepSFEnsemble
    meta, z = load('g1_reacher_ensemble')
    tr = gu.cuda_tr(transitions(z, 0, five=True))

    def build():
        sf = DeepSFEnsemble(pytorch_model_handle=gu.model_lambda(meta['hidden'], meta['acts']),
                            hyperparameters={'learning_rate_w': 0.5, 'learning_rate_sf': meta['lr']})
        sf.reset()
        for i in range(meta['N']):
            sf.add_training_task(gu.FakeTask(meta['S'], meta['A'], meta['D'], i))
        for i in range(meta['N']):
            layers = [(t(z[f'init.psi{i}.W{l}']), t(z[f'init.psi{i}.b{l}'])) for l in range(n_layers(meta))]
            gu.load_policy(sf, i, layers, t(z[f'init.w{i}']))
        return sf

    sf = build()                                                 # fused all-policy step == frozen-snapshot reference
    losses = sf.update_successors(tr)
    assert losses.shape == (meta['N'],)
    for i in range(meta['N']):
        for l, (W, b) in enumerate(gu.psi_params(sf, i)):
            assert rel_err(W, z[f'post_frozen.psi{i}.W{l}']) < STEP_TOL
            assert rel_err(b, z[f'post_frozen.psi{i}.b{l}']) < STEP_TOL
    sf = build()                                                 # literal reference loop (Gauss-Seidel), one call per task
    for i in range(meta['N']):
        sf.update_successor(tr, i)
    for i in range(meta['N']):
        for l, (W, b) in enumerate(gu.psi_params(sf, i)):
            assert rel_err(W, z[f'post_seq.psi{i}.W{l}']) < STEP_TOL


# ------------------------------------------------------------------------------------------------------------------
# oracle comparisons on fresh seeded inputs, at sizes the oracle finishes in seconds
# ------------------------------------------------------------------------------------------------------------------
def make_oracle(S, A, D, hidden, acts, N, seed, tsf_dim=None, beta=1):
    gen = torch.Generator().manual_seed(seed)
    o = OracleSF(S, A, D, hidden, acts, tsf_dim=tsf_dim, beta=beta)
    for _ in range(N):
        o.add_random_policy(gen)
    return o, gen


@pytest.mark.parametrize('S,A,D,hidden,N,B,hopper', [
    (4, 9, 12, (256, 256), 4, 4096, False),       # BASELINE config 1/2 shapes
    (11, 27, 50, (256, 256), 3, 777, True),       # Hopper shapes, ragged batch (not a multiple of the 64-row tile)
    (4, 2, 20, (256, 256), 3, 32, False),         # CartPole, config 5
    (7, 5, 9, (48, 48), 5, 130, False),           # odd everything: S, D, widths not multiples of 4/32
])
def test_forward_gpi_vs_oracle(S, A, D, hidden, N, B, hopper):
    meta = dict(S=S, A=A, D=D, hidden=list(hidden), acts=['relu'] * len(hidden), N=N)
    o, gen = make_oracle(S, A, D, hidden, meta['acts'], N, seed=77)
    sf = gu.build_g2(meta, oracle=o)
    x = synthetic_transitions(B, S, A, D, gen, hopper=hopper)[0]
    psi_ref = o.get_successors(x)
    assert rel_err(sf.get_successors(x.cuda()).cpu(), psi_ref) < FWD_TOL
    q_ref, task_ref = o.GPI(x, 1)
    q, task = sf.GPI(x.cuda(), 1)
    assert rel_err(q.cpu(), q_ref) < FWD_TOL
    ok, nbad = gu.argmax_mismatch_ok(q_ref, task_ref, task.cpu(), 'task', FWD_TOL)
    assert ok and nbad <= max(1, B // 1000)
    # action key: argmax_a max_j q (sfdqn.py:316)
    _, key_a, _ = sf._library.gpi(x.cuda(), sf.fit_w[1].weight, want_q=False)
    act, val = sf._library.decode_keys(key_a, want_value=True)
    act_ref = torch.argmax(torch.max(q_ref, dim=1).values, dim=-1)
    ok, nbad = gu.argmax_mismatch_ok(q_ref, act_ref, act.cpu(), 'action', FWD_TOL)
    assert ok and nbad <= max(1, B // 1000)
    assert rel_err(val.cpu(), q_ref.reshape(B, -1).max(dim=1).values) < FWD_TOL


@pytest.mark.parametrize('use_gpi', [True, False])
def test_g2_steps_vs_oracle_reacher_b4096(use_gpi):
    S, A, D, hidden, N, B = 4, 9, 12, (256, 256), 4, 4096
    meta = dict(S=S, A=A, D=D, hidden=list(hidden), acts=['relu', 'relu'], N=N)
    o, gen = make_oracle(S, A, D, hidden, meta['acts'], N, seed=5)
    sf = gu.build_g2(meta, oracle=o)
    for k in range(3):
        tr = synthetic_transitions(B, S, A, D, gen)
        ref = o.update_successor(tr, 2, use_gpi)
        out = sf.update_successor(gu.cuda_tr(tr), 2, use_gpi)
        assert np.allclose([float(v) for v in out], [float(v) for v in ref], rtol=2e-5, atol=1e-8)
    for l, (W, b) in enumerate(gu.psi_params(sf, 2)):
        assert rel_err(W, o.psi[2][l][0]) < STEP_TOL and mean_err(W, o.psi[2][l][0]) < 2e-6
        assert rel_err(b, o.psi[2][l][1]) < STEP_TOL
    assert rel_err(sf.fit_w[2].weight.data.cpu(), o.w[2]) < STEP_TOL


def test_g3_steps_vs_oracle_reacher_b4096():
    S, A, D, hidden, N, B = 4, 9, 12, (256, 256), 4, 4096
    meta = dict(S=S, A=A, D=D, hidden=list(hidden), acts=['relu', 'relu'], N=N, gdim=100, beta=1, use_gpi=True)
    o, gen = make_oracle(S, A, D, hidden, meta['acts'], N, seed=6, tsf_dim=100, beta=1)
    dsf, ag = gu.build_g3(meta, oracle=o)
    for k, pol in enumerate([0, 3, 0]):
        tr = synthetic_transitions(B, S, A, D, gen)
        ref = o.tsf_update_successor(tr, pol, True)
        out = ag.update_successor(gu.cuda_tr(tr), pol, True)
        assert np.allclose([float(v) for v in out], [float(v) for v in ref], rtol=2e-5, atol=1e-8)
    for pol in (0, 3):
        for l, (W, b) in enumerate(gu.psi_params(dsf, pol)):
            assert rel_err(W, o.psi[pol][l][0]) < STEP_TOL
        assert rel_err(ag.g_functions[pol].weight.data.cpu(), o.g[pol][0]) < STEP_TOL
    assert rel_err(ag.h_function.weight.data.cpu(), o.h[0]) < STEP_TOL
    assert rel_err(ag.h_function.bias.data.cpu(), o.h[1]) < STEP_TOL


@pytest.mark.parametrize('variant', ['g2', 'g3'])
def test_ensemble_all_policies_vs_frozen_oracle(variant):
    S, A, D, hidden, N, B = 4, 9, 12, (64, 64), 5, 200
    tsf = variant == 'g3'
    meta = dict(S=S, A=A, D=D, hidden=list(hidden), acts=['relu', 'relu'], N=N, gdim=16, beta=30, use_gpi=True)
    o, gen = make_oracle(S, A, D, hidden, meta['acts'], N, seed=9, tsf_dim=16 if tsf else None, beta=30)
    if tsf:
        sf, ag = gu.build_g3(meta, oracle=o)
    else:
        sf = ag = gu.build_g2(meta, oracle=o)
    for k in range(2):
        tr = synthetic_transitions(B, S, A, D, gen)
        ref = o.ensemble_update_frozen(tr, tsf=tsf, use_gpi=True)
        losses = ag.update_successor_all(gu.cuda_tr(tr), use_gpi=True).cpu()
        for i in range(N):
            assert np.allclose(losses[i].numpy(), [float(v) for v in ref[i]], rtol=3e-5, atol=1e-8)
    for i in range(N):
        for l, (W, b) in enumerate(gu.psi_params(sf, i)):
            assert rel_err(W, o.psi[i][l][0]) < STEP_TOL
        assert rel_err(sf.fit_w[i].weight.data.cpu(), o.w[i]) < STEP_TOL
    if tsf:
        assert rel_err(ag.h_function.weight.data.cpu(), o.h[0]) < STEP_TOL


def test_edge_cases_and_errors():
    meta = dict(S=4, A=9, D=12, hidden=[64, 64], acts=['relu', 'relu'], N=2)
    o, gen = make_oracle(4, 9, 12, (64, 64), meta['acts'], 2, seed=3)
    sf = gu.build_g2(meta, oracle=o)
    with pytest.raises(ValueError):
        sf.get_successors(torch.zeros(3, 5).cuda())              # wrong state width
    with pytest.raises(Exception):
        sf._library.train_step(gu.cuda_tr(synthetic_transitions(8, 4, 9, 12, gen)), 7)     # policy out of range
    assert sf.get_successors(torch.zeros(0, 4).cuda()).shape == (0, 2, 9, 12)             # empty batch
    # growing the library re-packs storage: earlier modules must still be live views
    for i in range(2, 7):
        sf.add_training_task(gu.FakeTask(4, 9, 12, i))
    x = torch.randn(5, 4, generator=gen)
    assert rel_err(sf.get_successor(x.cuda(), 1).cpu(), o.get_successor(x, 1)) < FWD_TOL
    W0 = gu.linears(sf.psi[0][0][0].net)[0].weight
    assert W0.data.data_ptr() == sf._library.online[0].data_ptr()
    # unsupported layer types are rejected, there is no fallback
    from deep_successor_features_for_transfer_b200.sfdqn import DeepSF
    bad = DeepSF(lambda *a: (torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Sigmoid(), torch.nn.Linear(8, 108)),
                             torch.nn.MSELoss(), None), hyperparameters=dict(gu.HYPER))
    bad.reset()
    with pytest.raises(TypeError):
        bad.add_training_task(gu.FakeTask(4, 9, 12, 0))


def test_full_size_properties_hopper_gpi():
    """BASELINE config 3 shapes at full batch (65 536 states), 8 policies = one GPU's shard: size-independent properties."""
    S, A, D, N, B = 11, 27, 50, 8, 65536
    meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N)
    o, gen = make_oracle(S, A, D, (256, 256), meta['acts'], N, seed=11)
    sf = gu.build_g2(meta, oracle=o)
    x = torch.sigmoid(torch.randn(B, S, generator=gen)).cuda()
    q, key_a, key_t = sf._library.gpi(x, sf.fit_w[0].weight)
    act, val = sf._library.decode_keys(key_a, want_value=True)
    task = sf._library.decode_keys(key_t)
    qmax = q.reshape(B, -1).max(dim=1).values
    assert torch.equal(val, qmax)                                                   # key value == max of the returned q
    assert torch.equal(act, torch.argmax(q.max(dim=1).values, dim=-1))              # first-index tie rule, actions
    assert torch.equal(task, torch.argmax(q.max(dim=2).values, dim=1))              # first-index tie rule, tasks
    # sharding invariance: max over two policy shards' keys == keys over all policies (what the NCCL MAX all-reduce does)
    _, ka0, kt0 = sf._library.gpi(x, sf.fit_w[0].weight, lo=0, n_pol=3, want_q=False, task_base=0)
    _, ka1, kt1 = sf._library.gpi(x, sf.fit_w[0].weight, lo=3, n_pol=5, want_q=False, task_base=3)
    assert torch.equal(torch.maximum(ka0, ka1), key_a) and torch.equal(torch.maximum(kt0, kt1), key_t)
    # linearity in w: q(2w) == 2 q(w) exactly (power-of-two scaling)
    q2, _, _ = sf._library.gpi(x, 2.0 * sf.fit_w[0].weight)
    assert torch.equal(q2, 2.0 * q)
    # spot-check 512 states against the oracle
    idx = torch.randperm(B, generator=gen)[:512]
    q_ref, _ = o.GPI(x.cpu()[idx], 0)
    assert rel_err(q.cpu()[idx], q_ref) < FWD_TOL


def test_device_replay_buffer_matches_host_ring():
    """the HBM ring + gather kernel return exactly what the reference-style host ring returns (same numpy stream)."""
    from deep_successor_features_for_transfer_b200.sfdqn import ReplayBuffer, DeviceReplayBuffer
    S, D, A, n_batch, cap = 4, 12, 9, 64, 300
    host, dev = ReplayBuffer(n_samples=cap, n_batch=n_batch), DeviceReplayBuffer(n_samples=cap, n_batch=n_batch)
    gen = torch.Generator().manual_seed(9)
    assert dev.replay() is None
    for k in range(cap + 57):                                     # wraps around the ring
        s, s1 = torch.randn(1, S, generator=gen), torch.randn(1, S, generator=gen)
        a = torch.randint(0, A, (), generator=gen)
        r, phi, gamma = float(torch.randn((), generator=gen)), torch.rand(D, generator=gen), 0.0 if k % 17 == 0 else 0.9
        host.append(s, a, r, phi, s1, gamma)
        dev.append(s.cuda() if k % 2 else s, a, r, phi, s1, gamma)      # host and device inputs both accepted
        if k == n_batch // 2:
            assert dev.replay() is None and host.replay() is None
    for seed in (1, 2):
        np.random.seed(seed)
        ref = host.replay()
        np.random.seed(seed)
        got = dev.replay()
        for name, x, y in zip(('states', 'actions', 'rewards', 'phis', 'next_states', 'gammas'), ref, got):
            assert x.shape == y.shape and x.dtype == y.dtype, (name, x.shape, y.shape, x.dtype, y.dtype)
            assert torch.equal(x.cpu(), y.cpu()), name


@pytest.mark.parametrize('fused', [True, False])
def test_g3_target_task_adaptation_vs_golden(fused):
    """TSFDQN.get_test_action / update_test_reward_mapper on the kernel-backed library == the reference: 5 adaptation
    steps of the unmodified reference (actions chosen, the three losses, w and omega after every step).  fused: the whole update
    incl. Adam(w, omega), the LambdaLR decay and the clamp is ONE kernel (csrc/target.cu) after the two ensemble forwards;
    fused=False: the eager op sequence with torch.optim (kept for callers that bring their own optimizer)."""
    meta, z = load('g3_reacher_target_adapt')
    dsf, ag = gu.build_g3(dict(meta, use_gpi=True), z=z)
    ag.hyperparameters.update({k: meta[k] for k in ('learning_rate_omega', 'weight_decay_omega', 'learning_rate_omega_decay',
                                                     'omegas_l1_coefficient')})
    ag.test_epsilon = 0.0
    ag.n_actions = meta['A']

    class Task:
        def __init__(self):
            self.k = 0

        def features(self, s, a, s1):
            return t(z[f'step{self.k}.phi'])

    task = Task()
    w_approx, optim, sched, omegas = ag._new_target_task(meta['D'], t(z['init.omegas']).cuda(), fused=fused)
    assert type(optim).__name__ == ('PackedTargetOptim' if fused else 'Adam')
    with torch.no_grad():
        w_approx.weight.copy_(t(z['init.w_target']))
    for k in range(meta['K']):
        task.k = k
        s, s1 = t(z[f'step{k}.s']).cuda(), t(z[f'step{k}.s1']).cuda()
        a, a1, r = int(z[f'step{k}.a']), int(z[f'step{k}.a1']), float(z[f'step{k}.r'])
        assert int(ag.get_test_action(s, w_approx, omegas)) == a and int(ag.get_test_action(s1, w_approx, omegas)) == a1
        losses = ag.update_test_reward_mapper(w_approx, omegas, optim, task, r, s, a, s1, a1)
        sched.step()
        assert np.allclose([float(x) for x in losses], z['out.losses'][k], rtol=2e-5, atol=1e-7)
        assert rel_err(omegas.detach().cpu(), z[f'step{k}.omegas']) < 2e-5
        assert rel_err(w_approx.weight.detach().cpu(), z[f'step{k}.w_target']) < 2e-5


# ---------------- learned features (PhiFunction + SFDQN.pre_train, sfdqn_phi.py:90-123, 800-873) ----------------
def _phi_fixture():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'phi_pretrain_cartpole.npz'))


def _seed_all(seed):
    import random
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def test_phi_function_forward_and_init_vs_golden():
    from deep_successor_features_for_transfer_b200.sfdqn_phi import PhiFunction
    z = _phi_fixture()
    S, D, seed = int(z['S']), int(z['D']), int(z['seed'])
    _seed_all(seed)
    phi = PhiFunction(S, 1, D)
    lin = [m for m in phi._model if isinstance(m, torch.nn.Linear)]
    for l, m in enumerate(lin):                                  # same draws as the reference's PhiFunction under the same seed
        assert np.array_equal(m.weight.data.cpu().numpy(), z[f'init.phi.W{l}'])
        assert np.array_equal(m.bias.data.cpu().numpy(), z[f'init.phi.b{l}'])
    with torch.no_grad():                                        # load the reference's trained weights through the module views
        for l, m in enumerate(lin):
            m.weight.data.copy_(torch.from_numpy(z[f'out.phi.W{l}']))
            m.bias.data.copy_(torch.from_numpy(z[f'out.phi.b{l}']))
    out = phi(torch.from_numpy(z['probe.s']), torch.from_numpy(z['probe.a']), torch.from_numpy(z['probe.s1']))
    assert out.shape == (16, D)
    assert np.allclose(out.cpu().numpy(), z['probe.phi'], rtol=1e-5, atol=1e-6)
    one = phi(torch.from_numpy(z['probe.s'][0]), torch.tensor(int(z['probe.a'][0])), torch.from_numpy(z['probe.s1'][0]))
    assert one.shape == (1, D) and np.allclose(one.cpu().numpy()[0], z['probe.phi'][0], rtol=1e-5, atol=1e-6)


def test_phi_regression_step_vs_oracle():
    """One update on a fixed batch: loss, dL/dtheta via the post-step weights, and both Adam states against OraclePhi."""
    from oracle.sf_oracle import OraclePhi
    from deep_successor_features_for_transfer_b200.sfdqn_phi import PhiFunction, _RewardHead
    z = _phi_fixture()
    S, A, D, seed = int(z['S']), int(z['A']), int(z['D']), int(z['seed'])
    _seed_all(seed)
    phi = PhiFunction(S, 1, D)
    head = _RewardHead(D, phi.device)
    o = OraclePhi([(torch.from_numpy(z[f'init.phi.W{l}']), torch.from_numpy(z[f'init.phi.b{l}'])) for l in range(3)])
    oh = dict(w=head.weight.cpu().reshape(1, D).clone(), m=torch.zeros(1, D), v=torch.zeros(1, D), step=0)
    g = torch.Generator().manual_seed(5)
    for B in (32, 300):                                          # 300: two CTAs of the head kernel (partials)
        s, a, s1 = torch.randn(B, S, generator=g), torch.randint(0, A, (B,), generator=g), torch.randn(B, S, generator=g)
        r = torch.tanh(torch.randn(B, 1, generator=g))
        for k in range(3):
            lo = o.regression_step(s, a, r, s1, oh)
            lg = float(phi.regression_step(s, a, r, s1, head))
            assert abs(lg - lo) <= 2e-5 * max(1.0, abs(lo))
    lin = [m for m in phi._model if isinstance(m, torch.nn.Linear)]
    for l, m in enumerate(lin):
        assert rel_err(m.weight.data.cpu(), o.layers[l][0]) < 2e-5
        assert rel_err(m.bias.data.cpu(), o.layers[l][1]) < 2e-5
    assert rel_err(head.weight.cpu().reshape(1, D), oh['w']) < 2e-5
    assert int(head.step) == oh['step'] == 6 and int(phi._library.step[0]) == 6


def test_phi_pre_train_vs_golden():
    """The whole pre_train loop under the reference's seeds: same replay picks, same random actions, losses along the way."""
    from deep_successor_features_for_transfer_b200.sfdqn_phi import pre_train
    from tests.phi_util import FakePhiTask
    z = _phi_fixture()
    S, A, D, n_tasks, seed = (int(z[k]) for k in ('S', 'A', 'D', 'n_tasks', 'seed'))
    _seed_all(seed)
    tasks = [FakePhiTask(S, A, D, i) for i in range(n_tasks)]
    model, losses = pre_train(tasks, int(z['n_samples']), int(z['n_cycles']))
    ref = z['out.losses']
    assert len(losses) == len(ref)
    assert np.allclose(losses[:20], ref[:20], rtol=1e-4, atol=1e-7)
    assert np.allclose(losses, ref, rtol=2e-2, atol=1e-4)           # 149 chained Adam updates: fp32 summation-order drift
    out = model(torch.from_numpy(z['probe.s']), torch.from_numpy(z['probe.a']), torch.from_numpy(z['probe.s1']))
    assert rel_err(out.cpu(), torch.from_numpy(z['probe.phi'])) < 2e-2


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_add_policy_without_growth_drops_cached_plans(precision):
    """
    The reference's G1 flow adds a task, trains, adds the next (agents/agent.py train_on_task; agents/sfdqn.py:59 steps every
    index each batch).  A library of capacity 4 grows 2 -> 3 policies WITHOUT reallocating: the cached train-step plans (GPI
    range, pack range, 'all' = n policies) must not survive add_policy.  After the third policy arrives, the single-policy
    step's GPI must see all 3 nets and the all-task step must train all 3 -- checked against the oracle.
    """
    S, A, D, hidden, B = 4, 9, 12, (256, 256), 384
    tol_l, tol_w = (3e-5, STEP_TOL) if precision == 'fp32' else (3e-2, None)
    meta = dict(S=S, A=A, D=D, hidden=list(hidden), acts=['relu', 'relu'], N=2)
    o, gen = make_oracle(S, A, D, hidden, meta['acts'], 3, seed=13)
    third = (o.psi.pop(), o.tgt.pop(), o.w.pop(), o.adam.pop(), o.updates_since_target_updated.pop())
    sf = gu.build_g2(meta, oracle=o, hyper=dict(gu.HYPER, precision=precision))
    assert sf._library.cap == 4
    tr = [synthetic_transitions(B, S, A, D, gen) for _ in range(4)]
    # plans built at n = 2
    for k, fn in ((0, lambda b: sf.update_successor(b, 0, True)), (1, lambda b: sf.update_successor_all(b, use_gpi=True))):
        ref = o.update_successor(tr[k], 0, True) if k == 0 else o.ensemble_update_frozen(tr[k], use_gpi=True)
        got = fn(gu.cuda_tr(tr[k]))
        got = [float(v) for v in got] if k == 0 else got.cpu().tolist()[0]
        assert np.allclose(got, [float(v) for v in (ref if k == 0 else ref[0])], rtol=tol_l, atol=1e-8)
    # third policy: no capacity growth
    sf.add_training_task(gu.FakeTask(S, A, D, 2))
    assert sf._library.cap == 4 and sf._library.n == 3
    gu.load_policy(sf, 2, third[0], third[2])
    o.psi.append(third[0]); o.tgt.append(third[1]); o.w.append(third[2]); o.adam.append(third[3]); o.updates_since_target_updated.append(third[4])
    # make the newcomer dominate GPI so that a stale 2-policy plan would give different next actions
    with torch.no_grad():
        big = o.psi[2][-1][1] + 5.0 * torch.sign(o.w[0].reshape(-1)).repeat(A)
        o.psi[2][-1] = (o.psi[2][-1][0], big)
        gu.linears(sf.psi[2][0][0].net)[-1].bias.data.copy_(big)
    ref = o.update_successor(tr[2], 0, True)
    got = sf.update_successor(gu.cuda_tr(tr[2]), 0, True)
    assert np.allclose([float(v) for v in got], [float(v) for v in ref], rtol=tol_l, atol=1e-8)
    ref = o.ensemble_update_frozen(tr[3], use_gpi=True)
    got = sf.update_successor_all(gu.cuda_tr(tr[3]), use_gpi=True).cpu()
    assert got.shape[0] == 3
    for i in range(3):
        assert np.allclose(got[i].numpy(), [float(v) for v in ref[i]], rtol=tol_l, atol=1e-8)
    if tol_w is not None:
        for i in range(3):
            for l, (W, b) in enumerate(gu.psi_params(sf, i)):
                assert rel_err(W, o.psi[i][l][0]) < tol_w
    assert int(sf._library.step[2]) == 1                      # the newcomer was stepped exactly once (the 'all' plan saw it)


def test_g4_joint_psi_phi_vs_golden():
    """
    G4 joint psi / phi step on the kernels (DeepSF_PHI.update_successor, features/deep_phi.py:95-224 + the phi model of
    main_sfdqn_phi_torch.py:52-73 + the loss coefficient of agents/sfdqn_phi.py:152-165) against the UNMODIFIED reference's
    4 recorded steps (tests/golden/make_golden_g4.py): (loss, psi_loss, phi_loss, coefficient) per step at 2e-5, post-step psi /
    phi / fit_w incl. bias at STEP_TOL, untouched policies bit-equal.  Every update is the first step of a fresh Adam
    (p -= lr * g / (|g| + eps)), the coefficient's group has maximize=True and is clamped to [1e-2, 1e6].
    """
    import os
    from deep_successor_features_for_transfer_b200.sfdqn_phi import DeepSF_PHI, PackedPhi
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'g4_joint_psi_phi.npz'))
    S, A, D, N, K, policy = (int(z[k]) for k in ('S', 'A', 'D', 'N', 'K', 'policy'))
    hidden, acts = [int(h) for h in z['hidden']], [str(a) for a in z['acts']]
    sf = DeepSF_PHI(pytorch_model_handle=gu.model_lambda(hidden, acts), target_update_ev=1000, hyperparameters=dict(gu.HYPER))
    sf.reset()
    for i in range(N):
        sf.add_training_task(gu.FakeTask(S, A, D, i))
    n_psi = len(hidden) + 2
    with torch.no_grad():
        for i in range(N):
            layers = [(t(z[f'init.psi{i}.W{l}']), t(z[f'init.psi{i}.b{l}'])) for l in range(n_psi)]
            gu.load_policy(sf, i, layers, t(z[f'init.w{i}.W']))
            sf.fit_w[i].bias.data.copy_(t(z[f'init.w{i}.b']))
    n_in = 2 * S + 1
    phi_net = torch.nn.Sequential(torch.nn.Linear(n_in, 2 * n_in), torch.nn.ReLU(), torch.nn.Linear(2 * n_in, 2 * n_in), torch.nn.ReLU(),
                                  torch.nn.Linear(2 * n_in, 2 * n_in), torch.nn.ReLU(), torch.nn.Linear(2 * n_in, 2 * n_in), torch.nn.ReLU(),
                                  torch.nn.Linear(2 * n_in, D))
    phi = PackedPhi(phi_net, D)
    with torch.no_grad():
        for l, lin in enumerate(gu.linears(phi.net)):
            lin.weight.data.copy_(t(z[f'init.phi.W{l}']))
            lin.bias.data.copy_(t(z[f'init.phi.b{l}']))
    phis_model = ((phi, torch.nn.MSELoss(), None), None)
    coef = torch.ones(1, device='cuda')
    untouched = [W.clone() for W, _ in gu.psi_params(sf, 0)]
    for k in range(K):
        tr = tuple(t(z[f'tr{k}.{n_}']) for n_ in ('states', 'actions', 'rs', 'phis', 'next_states', 'gammas'))
        tr = (tr[0], tr[1].long(), tr[2], tr[3], tr[4], tr[5])
        out = sf.update_successor(gu.cuda_tr(tr), phis_model, policy, coef, bool(int(z['use_gpi'])))
        got = [float(v) for v in out]
        assert np.allclose(got, z['out.losses'][k], rtol=2e-5, atol=1e-7), (k, got, z['out.losses'][k])
    for l, (W, b) in enumerate(gu.psi_params(sf, policy)):
        assert rel_err(W, z[f'post.psi.W{l}']) < STEP_TOL and rel_err(b, z[f'post.psi.b{l}']) < STEP_TOL
    for l, lin in enumerate(gu.linears(phi.net)):
        assert rel_err(lin.weight.data.cpu(), z[f'post.phi.W{l}']) < STEP_TOL
        assert rel_err(lin.bias.data.cpu(), z[f'post.phi.b{l}']) < STEP_TOL
    assert rel_err(sf.fit_w[policy].weight.data.cpu(), z['post.w.W']) < STEP_TOL
    assert rel_err(sf.fit_w[policy].bias.data.cpu(), z['post.w.b']) < STEP_TOL
    assert all(torch.equal(a, b) for a, (b, _) in zip(untouched, gu.psi_params(sf, 0)))
    # q of GPI includes fit_w's bias (w(psi), deep_phi.py:248)
    x = t(z['tr0.states']).cuda()
    q, _ = sf.GPI(x, policy)
    psi = sf.get_successors(x)
    q_ref = torch.nn.functional.linear(psi, sf.fit_w[policy].weight.data, sf.fit_w[policy].bias.data)[..., 0]
    assert rel_err(q.cpu(), q_ref.cpu()) < 1e-5


def test_g1_lms_update_reward_kernel():
    """G1's per-step LMS rule (features/successor.py:146-167) as one kernel == the torch formula, on the packed reward row."""
    from deep_successor_features_for_transfer_b200.ensemble import DeepSF as DeepSF_G1
    sf = DeepSF_G1(pytorch_model_handle=gu.model_lambda([64, 64], ['relu', 'relu']), hyperparameters={'learning_rate_w': 0.5})
    sf.reset()
    for i in range(3):
        sf.add_training_task(gu.FakeTask(4, 9, 12, i))
    gen = torch.Generator().manual_seed(3)
    w_ref = sf.fit_w[1].clone().cpu()
    for _ in range(20):
        phi, r = torch.rand(12, generator=gen) * 2.5 - 1.5, torch.randn((), generator=gen)
        w_ref = w_ref + 0.5 * (r - torch.sum(phi.reshape(-1, 1) * w_ref)) * phi.reshape(-1, 1)
        sf.update_reward(phi, r, 1)
    assert rel_err(sf.fit_w[1].cpu(), w_ref) < 1e-5
    assert torch.equal(sf.fit_w[1], sf._library.w[1].view(-1, 1))          # in place on the packed row


def _build_nf(meta, g_of, h, psi_of, w_of, precision='fp32'):
    """TSF agent with planar-flow g functions (deep_successor_features_for_transfer_b200.tsfdqn_nf) loaded with given CPU tensors."""
    from deep_successor_features_for_transfer_b200.tsfdqn_nf import DeepTSF, TSFDQN, ReplayBuffer
    hyper = dict(gu.HYPER, g_h_function_dims=meta['gdim'], beta_loss_coefficient=meta['beta'], n_coupling_layers=meta['n_flows'],
                 precision=precision)
    dsf = DeepTSF(pytorch_model_handle=gu.model_lambda(meta['hidden'], meta['acts']), use_true_reward=False, target_update_ev=1000,
                  hyperparameters=hyper)
    ag = TSFDQN(deep_sf=dsf, buffer_handle=lambda: ReplayBuffer(), gamma=0.9, T=500, encoding=None, use_gpi=True, hyperparameters=hyper)
    ag.reset()
    for i in range(meta['N']):
        ag.add_training_task(gu.FakeTask(meta['S'], meta['A'], meta['D'], i))
    with torch.no_grad():
        for i in range(meta['N']):
            gu.load_policy(dsf, i, psi_of(i), w_of(i))
            mods = list(ag.g_functions[i])
            assert len(mods) == meta['n_flows'] + 1 and sum(1 for _ in ag.g_functions[i].parameters()) == 3 * meta['n_flows'] + 2
            for f, (fw, fb, fs) in zip(mods[:-1], g_of(i)[:-1]):
                f.weight.data.copy_(fw); f.bias.data.copy_(fb); f.scale.data.copy_(fs)
            mods[-1].weight.data.copy_(g_of(i)[-1][0]); mods[-1].bias.data.copy_(g_of(i)[-1][1])
        ag.h_function.weight.data.copy_(h[0]); ag.h_function.bias.data.copy_(h[1])
    return dsf, ag


def test_g3_normalising_flow_g_vs_golden():
    """tsfdqn_nf.py (PlanarFlow :331-358, update :620-720) on the kernels: the unmodified reference's 4 steps with 5 planar flows
    per g -- losses, post-step psi / w / h and every flow parameter."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'g3_nf_reacher.npz'))
    meta = dict(S=int(z['S']), A=int(z['A']), D=int(z['D']), N=int(z['N']), gdim=int(z['gdim']), beta=int(z['beta']),
                n_flows=int(z['n_flows']), hidden=[int(v) for v in z['hidden']], acts=[str(a) for a in z['acts']])
    K, policy, nf = int(z['K']), int(z['policy']), meta['n_flows']
    n_psi = len(meta['hidden']) + 2
    g_of = lambda i: [(t(z[f'init.g{i}.f{k}.weight']), t(z[f'init.g{i}.f{k}.bias']), t(z[f'init.g{i}.f{k}.scale'])) for k in range(nf)] + \
        [(t(z[f'init.g{i}.W']), t(z[f'init.g{i}.b']))]
    dsf, ag = _build_nf(meta, g_of, (t(z['init.h.W']), t(z['init.h.b'])),
                        lambda i: [(t(z[f'init.psi{i}.W{l}']), t(z[f'init.psi{i}.b{l}'])) for l in range(n_psi)], lambda i: t(z[f'init.w{i}']))
    for k in range(K):
        tr = tuple(t(z[f'tr{k}.{n}']) for n in ('states', 'actions', 'rs', 'phis', 'next_states', 'gammas'))
        out = ag.update_successor(gu.cuda_tr(tr), policy, True)
        assert np.allclose([float(v) for v in out], z['out.losses'][k], rtol=2e-5, atol=1e-7), (k, out, z['out.losses'][k])
    for l, (W, b) in enumerate(gu.psi_params(dsf, policy)):
        assert np.allclose(W.numpy(), z[f'post.psi.W{l}'], rtol=1e-4, atol=2e-6)
    mods = list(ag.g_functions[policy])
    for k, f in enumerate(mods[:-1]):
        for name in ('weight', 'bias', 'scale'):
            assert np.allclose(getattr(f, name).data.cpu().numpy(), z[f'post.g.f{k}.{name}'], rtol=1e-4, atol=2e-6), (k, name)
    assert np.allclose(mods[-1].weight.data.cpu().numpy(), z['post.g.W'], rtol=1e-4, atol=2e-6)
    assert np.allclose(mods[-1].bias.data.cpu().numpy(), z['post.g.b'], rtol=1e-4, atol=2e-6)
    assert np.allclose(ag.h_function.weight.data.cpu().numpy(), z['post.h.W'], rtol=1e-4, atol=2e-6)
    assert np.allclose(dsf.fit_w[policy].weight.data.cpu().numpy(), z['post.w'], rtol=1e-4, atol=2e-6)
    for i in range(meta['N']):                                   # the other policies' flows are untouched
        if i != policy:
            assert torch.equal(list(ag.g_functions[i])[0].weight.data.cpu(), g_of(i)[0][0])


@pytest.mark.parametrize('precision,n_flows', [('fp32', 3), ('bf16', 3), ('fp32', 10)])
def test_g3_normalising_flow_steps_vs_oracle(precision, n_flows):
    """Planar-flow g at the bench's shape (256-wide psi nets, B = 1000: several thread-block clusters, a ragged last tile) against
    the oracle: single-policy steps and the all-task ensemble step; flows, Linear and h within the fp32 path's 1e-4 (the flows'
    arithmetic is fp32 in every precision mode; bf16 mode: psi-dependent quantities at that mode's bounds)."""
    from oracle.sf_oracle import init_linear
    S, A, D, N, B, G = 4, 9, 12, 3, 1000, 100
    meta = dict(S=S, A=A, D=D, N=N, gdim=G, beta=30, n_flows=n_flows, hidden=[256, 256], acts=['relu', 'relu'])
    gen = torch.Generator().manual_seed(17)
    o = OracleSF(S, A, D, (256, 256), ('relu', 'relu'), tsf_dim=G, beta=30)
    h = init_linear(D, G, gen)
    for i in range(N):
        psi = [init_linear(256, S, gen), init_linear(256, 256, gen), init_linear(256, 256, gen), init_linear(A * D, 256, gen)]
        flows = [((torch.rand(1, S, generator=gen) - 0.5) * 0.6, (torch.rand(1, generator=gen) - 0.5) * 0.6,
                  (torch.rand(1, S, generator=gen) - 0.5) * 0.6) for _ in range(n_flows)]       # (larger than the init range: tanh is exercised)
        o.add_policy(psi, (torch.rand(1, D, generator=gen) * 0.02 - 0.01), flows + [init_linear(G, S, gen)], h)
    dsf, ag = _build_nf(meta, lambda i: o.g[i], o.h, lambda i: o.psi[i], lambda i: o.w[i], precision)
    ltol = 2e-5 if precision == 'fp32' else 3e-2
    for pol in (1, 2, 1):
        tr = synthetic_transitions(B, S, A, D, gen)
        ref = o.tsf_update_successor(tr, pol, True)
        out = ag.update_successor(gu.cuda_tr(tr), pol, True)
        assert np.allclose([float(v) for v in out], [float(v) for v in ref], rtol=ltol, atol=1e-7), (pol, out, ref)
    if precision == 'fp32':
        for pol in (1, 2):
            mods = list(ag.g_functions[pol])
            for k, f in enumerate(mods[:-1]):
                for j, name in enumerate(('weight', 'bias', 'scale')):
                    assert rel_err(getattr(f, name).data.cpu(), o.g[pol][k][j]) < 1e-4, (pol, k, name)
            assert rel_err(mods[-1].weight.data.cpu(), o.g[pol][-1][0]) < 1e-4
        assert rel_err(ag.h_function.weight.data.cpu(), o.h[0]) < 1e-4
    tr = synthetic_transitions(B, S, A, D, gen)
    ref = o.ensemble_update_frozen(tr, tsf=True, use_gpi=True)
    losses = ag.update_successor_all(gu.cuda_tr(tr), use_gpi=True).cpu()
    for i in range(N):
        assert np.allclose(losses[i].numpy(), [float(v) for v in ref[i]], rtol=ltol, atol=1e-7), (i, losses[i], ref[i])
