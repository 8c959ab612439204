# -*- coding: UTF-8 -*-
"""
Agent-level golden vectors: the UNMODIFIED reference agents (sfdqn.SFDQN / tsfdqn.TSFDQN) run their own
next_sample -> train_agent -> ReplayBuffer.replay -> update_successor loop (sfdqn.py:462-484, 550-627; tsfdqn.py:435-497,
566-586, 588-709) on the deterministic ToyTask environment under fixed `random` / `numpy` / `torch` seeds.  The loop driven
here is the body of `train()` (sfdqn.py:629-677) without the target-task evaluation: reset, add_training_task per task, then
per task set_active_training_task + n_samples x (next_sample; total_training_steps += 1).  (train() itself re-draws the
initial weights inside; driving its body lets the fixture record them.)

Recorded per environment step: the action taken, the replay picks' effect (the three losses; NaN while the ring is still
filling), and at the end the stepped policy's weights, w, (g, h), the GPI counters and epsilon.

    python tests/golden/make_golden_agent.py
"""
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                            # noqa: E402  (reference import recipe, model lambda, dump helpers)

from tests.toy_task import ToyTask                                  # noqa: E402

HIDDEN, ACTS = (64, 64), ('relu', 'relu')


def run_agent(kind, name, S, A, D, N, n_samples, n_batch, T, epsilon, use_gpi, seed, beta=1, gdim=100):
    torch.manual_seed(seed)
    hyper = dict(mg.HYPER, g_h_function_dims=gdim, beta_loss_coefficient=beta)
    ref = mg.ref_sfdqn if kind == 'sfdqn' else mg.ref_tsfdqn
    with contextlib.redirect_stdout(io.StringIO()):
        if kind == 'sfdqn':
            sf = ref.DeepSF(pytorch_model_handle=mg.make_model_lambda(HIDDEN, ACTS), use_true_reward=False,
                            target_update_ev=25, hyperparameters=hyper)
            ag = ref.SFDQN(deep_sf=sf, buffer_handle=lambda: ref.ReplayBuffer(n_batch=n_batch), gamma=0.9, T=T, encoding=None,
                           epsilon=epsilon, use_gpi=use_gpi, hyperparameters=hyper)
        else:
            sf = ref.DeepTSF(pytorch_model_handle=mg.make_model_lambda(HIDDEN, ACTS), use_true_reward=False,
                             target_update_ev=25, hyperparameters=hyper)
            ag = ref.TSFDQN(deep_sf=sf, buffer_handle=lambda: ref.ReplayBuffer(n_batch=n_batch), gamma=0.9, T=T, encoding=None,
                            epsilon=epsilon, use_gpi=use_gpi, hyperparameters=hyper)
        tasks = [ToyTask(S, A, D, i, seed=seed) for i in range(N)]
        ag.reset()
        for task in tasks:
            ag.add_training_task(task)
    out = {'meta': np.array(repr(dict(kind='agent_' + kind, S=S, A=A, D=D, hidden=list(HIDDEN), acts=list(ACTS), N=N,
                                       n_samples=n_samples, n_batch=n_batch, T=T, epsilon=epsilon, use_gpi=use_gpi, seed=seed,
                                       beta=beta, gdim=gdim, target_update_ev=25, gamma=0.9)))}
    for i in range(N):
        mg.dump_net(out, f'init.psi{i}', sf.psi[i][0][0])
        out[f'init.w{i}'] = sf.fit_w[i].weight.detach().numpy().copy()
        if kind == 'tsfdqn':
            out[f'init.g{i}.W'] = ag.g_functions[i].weight.detach().numpy().copy()
            out[f'init.g{i}.b'] = ag.g_functions[i].bias.detach().numpy().copy()
    if kind == 'tsfdqn':
        out['init.h.W'] = ag.h_function.weight.detach().numpy().copy()
        out['init.h.b'] = ag.h_function.bias.detach().numpy().copy()

    # record every update's losses without touching the reference's code: wrap the bound method on the instance
    losses = []
    holder = sf if kind == 'sfdqn' else ag
    inner = holder.update_successor

    def recording(transitions, policy_index, use_gpi=True):
        res = inner(transitions, policy_index, use_gpi)
        losses.append([float('nan')] * 3 if res is None else [float(x) for x in res])
        return res
    holder.update_successor = recording

    random.seed(seed)
    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        for index in range(N):
            ag.set_active_training_task(index)
            for t in range(n_samples):
                ag.next_sample(None, None)
                ag.total_training_steps += 1
    out['out.actions'] = np.array([a for task in tasks for a in task.actions_taken], dtype=np.int64)
    out['out.losses'] = np.array(losses, dtype=np.float64)
    out['out.epsilon'] = np.array(ag.epsilon)
    out['out.cum_reward'] = np.array(float(ag.cum_reward))
    for i in range(N):
        out[f'out.gpi_counters{i}'] = np.asarray(sf.gpi_counters[i]).astype(np.int64)
        mg.dump_net(out, f'post.psi{i}', sf.psi[i][0][0])
        mg.dump_net(out, f'post.tgt{i}', sf.psi[i][1][0])
        out[f'post.w{i}'] = sf.fit_w[i].weight.detach().numpy().copy()
        if kind == 'tsfdqn':
            out[f'post.g{i}.W'] = ag.g_functions[i].weight.detach().numpy().copy()
            out[f'post.g{i}.b'] = ag.g_functions[i].bias.detach().numpy().copy()
    if kind == 'tsfdqn':
        out['post.h.W'] = ag.h_function.weight.detach().numpy().copy()
        out['post.h.b'] = ag.h_function.bias.detach().numpy().copy()
    out['post.updates_since_target_updated'] = np.array(sf.updates_since_target_updated)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    n_upd = int(np.isfinite(out['out.losses'][:, 0]).sum())
    print('wrote', name, 'steps', len(out['out.actions']), 'updates', n_upd, 'last losses', losses[-1],
          'gpi counters', [out[f'out.gpi_counters{i}'].tolist() for i in range(N)])


if __name__ == '__main__':
    run_agent('sfdqn', 'agent_sfdqn_toy', S=4, A=5, D=6, N=2, n_samples=100, n_batch=8, T=20, epsilon=0.25, use_gpi=True, seed=11)
    run_agent('tsfdqn', 'agent_tsfdqn_toy', S=4, A=5, D=6, N=2, n_samples=100, n_batch=8, T=20, epsilon=0.25, use_gpi=True,
              seed=12, beta=3)
    run_agent('sfdqn', 'agent_sfdqn_nogpi_toy', S=4, A=5, D=6, N=2, n_samples=60, n_batch=8, T=20, epsilon=0.25, use_gpi=False,
              seed=13)
