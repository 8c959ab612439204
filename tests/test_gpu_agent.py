"""
-m gpu: AGENT-level drop-in parity.  The CUDA agents (deep_successor_features_for_transfer_b200.sfdqn.SFDQN / tsfdqn.TSFDQN)
run the reference's own loop -- set_active_training_task, then next_sample (GPI action selection from the packed keys,
environment step, train_agent -> buffer.append -> buffer.replay -> update_successor) -- on the deterministic ToyTask under
the same `random` / `numpy` seeds as the UNMODIFIED reference agents did when tests/golden/make_golden_agent.py recorded
them.  Checked over the whole run: every action taken (exact), every loss (1e-3: the trajectories are identical, fp32
summation order drifts over up to 186 Adam steps), the GPI counters (exact), final weights (1e-3), target-sync bookkeeping.
Also: the same run with the device-resident replay ring, and the batch-1 selection call on its own.
"""
import random

import numpy as np
import pytest
import torch

from tests import gpu_util as gu
from tests.golden_util import load, n_layers, rel_err, t
from tests.toy_task import ToyTask

pytestmark = pytest.mark.gpu


def build_agent(meta, z, device_ring=False):
    from deep_successor_features_for_transfer_b200 import sfdqn as m_sf, tsfdqn as m_tsf
    tsf = meta['kind'] == 'agent_tsfdqn'
    hyper = dict(gu.HYPER, g_h_function_dims=meta['gdim'], beta_loss_coefficient=meta['beta'])
    handle = gu.model_lambda(meta['hidden'], meta['acts'])
    ring = (lambda: m_sf.DeviceReplayBuffer(n_samples=4096, n_batch=meta['n_batch'])) if device_ring else \
        (lambda: m_sf.ReplayBuffer(n_batch=meta['n_batch']))
    if tsf:
        sf = m_tsf.DeepTSF(pytorch_model_handle=handle, target_update_ev=meta['target_update_ev'], hyperparameters=hyper)
        ag = m_tsf.TSFDQN(deep_sf=sf, buffer_handle=ring, gamma=meta['gamma'], T=meta['T'], encoding=None,
                          epsilon=meta['epsilon'], use_gpi=meta['use_gpi'], hyperparameters=hyper)
    else:
        sf = m_sf.DeepSF(pytorch_model_handle=handle, target_update_ev=meta['target_update_ev'], hyperparameters=hyper)
        ag = m_sf.SFDQN(deep_sf=sf, buffer_handle=ring, gamma=meta['gamma'], T=meta['T'], encoding=None,
                        epsilon=meta['epsilon'], use_gpi=meta['use_gpi'], hyperparameters=hyper)
    tasks = [ToyTask(meta['S'], meta['A'], meta['D'], i, seed=meta['seed']) for i in range(meta['N'])]
    ag.reset()
    for task in tasks:
        ag.add_training_task(task)
    with torch.no_grad():
        for i in range(meta['N']):
            layers = [(t(z[f'init.psi{i}.W{l}']), t(z[f'init.psi{i}.b{l}'])) for l in range(n_layers(meta))]
            gu.load_policy(sf, i, layers, t(z[f'init.w{i}']))
            if tsf:
                ag.g_functions[i].weight.data.copy_(t(z[f'init.g{i}.W']))
                ag.g_functions[i].bias.data.copy_(t(z[f'init.g{i}.b']))
        if tsf:
            ag.h_function.weight.data.copy_(t(z['init.h.W']))
            ag.h_function.bias.data.copy_(t(z['init.h.b']))
    return sf, ag, tasks


def drive(meta, sf, ag, tasks):
    """The body of train() (sfdqn.py:663-676) without target-task evaluation; records every update's losses."""
    holder = ag if meta['kind'] == 'agent_tsfdqn' else sf
    inner, losses = holder.update_successor, []

    def recording(transitions, policy_index, use_gpi=True):
        res = inner(transitions, policy_index, use_gpi)
        losses.append(None if res is None else torch.stack(list(res)))
        return res
    holder.update_successor = recording
    random.seed(meta['seed'])
    np.random.seed(meta['seed'])
    for index in range(meta['N']):
        ag.set_active_training_task(index)
        for _ in range(meta['n_samples']):
            ag.next_sample(None, None)
            ag.total_training_steps += 1
    nan = [float('nan')] * 3
    return np.array([nan if l is None else l.cpu().tolist() for l in losses])


@pytest.mark.parametrize('name,device_ring', [('agent_sfdqn_toy', False), ('agent_tsfdqn_toy', False),
                                              ('agent_sfdqn_nogpi_toy', False), ('agent_sfdqn_toy', True),
                                              ('agent_tsfdqn_toy', True)])
def test_agent_train_loop_vs_reference(name, device_ring):
    meta, z = load(name)
    sf, ag, tasks = build_agent(meta, z, device_ring)
    losses = drive(meta, sf, ag, tasks)
    actions = np.array([a for task in tasks for a in task.actions_taken], dtype=np.int64)
    ref_a, ref_l = z['out.actions'], z['out.losses']
    first_bad = int(np.argmax(actions != ref_a)) if not np.array_equal(actions, ref_a) else -1
    assert first_bad < 0, f'action sequence leaves the reference at environment step {first_bad}'
    assert np.array_equal(np.isnan(losses), np.isnan(ref_l))
    ok = ~np.isnan(ref_l)
    assert np.allclose(losses[ok], ref_l[ok], rtol=1e-3, atol=1e-5)
    for i in range(meta['N']):
        assert np.array_equal(np.asarray(sf.gpi_counters[i]), z[f'out.gpi_counters{i}'])
        for l, (W, b) in enumerate(gu.psi_params(sf, i)):
            assert rel_err(W, z[f'post.psi{i}.W{l}']) < 1e-3 and rel_err(b, z[f'post.psi{i}.b{l}']) < 1e-3
        for l, (W, b) in enumerate(gu.psi_params(sf, i, target=True)):
            assert rel_err(W, z[f'post.tgt{i}.W{l}']) < 1e-3
        assert rel_err(sf.fit_w[i].weight.data.cpu(), z[f'post.w{i}']) < 1e-3
        if meta['kind'] == 'agent_tsfdqn':
            assert rel_err(ag.g_functions[i].weight.data.cpu(), z[f'post.g{i}.W']) < 1e-3
    if meta['kind'] == 'agent_tsfdqn':
        assert rel_err(ag.h_function.weight.data.cpu(), z['post.h.W']) < 1e-3
    assert list(sf.updates_since_target_updated) == list(z['post.updates_since_target_updated'])
    assert abs(float(ag.cum_reward) - float(z['out.cum_reward'])) < 1e-3 * abs(float(z['out.cum_reward']))


def test_greedy_action_equals_reference_selection():
    """A9 (sfdqn.py:585-594): the packed-key selection equals q[:, c, :].argmax() of the reference-shaped GPI call, batch 1."""
    meta, z = load('agent_sfdqn_toy')
    sf, ag, tasks = build_agent(meta, z)
    gen = torch.Generator().manual_seed(3)
    for k in range(50):
        s = torch.randn(1, meta['S'], generator=gen)
        for i in range(meta['N']):
            q, c = sf.GPI(s.cuda(), i)
            a = sf.greedy_action(s, i, use_gpi=True)
            assert int(a) == int(torch.argmax(q[:, int(c), :].flatten()))
            a = sf.greedy_action(s, i, use_gpi=False)
            assert int(a) == int(torch.argmax(q[:, i, :].flatten()))
    counts = [int(np.sum(c)) for c in sf.gpi_counters]
    assert counts == [50, 50]                                  # counted once per use_gpi selection, as GPI(update_counters=True)
