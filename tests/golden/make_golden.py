# -*- coding: UTF-8 -*-
"""
Generates tests/golden/*.npz by executing the UNMODIFIED reference (okgarces/deep-successor-features-for-transfer,
mounted read-only at /root/reference) on seeded synthetic inputs.  Run once in the build container:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, so the vectors are committed; this script is the provenance.
Import recipe: SURVEY.md appendix A (stub matplotlib, set the device singleton BEFORE importing sfdqn/tsfdqn).
Each fixture stores: the case description, the initial weights of every reference module, the transitions fed, and the
reference's outputs (psi, q, task, losses per step, post-step weights, Adam exp_avg / exp_avg_sq / step, targets).
"""
import contextlib
import copy
import io
import os
import sys
import types
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference/source")
from utils.torch import set_torch_device, get_activation          # noqa: E402
from utils.logger import set_logger_level                         # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    set_torch_device(use_gpu=False)
    set_logger_level(use_logger=False)
import sfdqn as ref_sfdqn                                          # noqa: E402
import tsfdqn as ref_tsfdqn                                        # noqa: E402
from features.deep import DeepSF as RefDeepSF_G1                   # noqa: E402

from oracle.sf_oracle import synthetic_transitions                 # noqa: E402  (input generator only)

torch.set_num_threads(1)


class FakeTask:
    """Task protocol consumed by the SF library (tasks/task.py): shapes only."""

    def __init__(self, S, A, D, index):
        self.S, self.A, self.D, self.index = S, A, D, index

    def action_count(self):
        return self.A

    def feature_dim(self):
        return self.D

    def encode_dim(self):
        return self.S

    def get_w(self):
        w = torch.zeros(self.D, 1)
        w[self.index % self.D, 0] = 1.0
        return w

    def features(self, s, a, s1):
        return torch.zeros(self.D)


def make_model_lambda(n_neurons, activations, with_optim_lr=None):
    """The mains' sf_model_lambda (main_tsfdqn_sequential_torch.py:44-75), parameterised instead of reading a cfg."""

    def sf_model_lambda(num_inputs, output_dim, reshape_dim, reshape_axis=1):
        layers = OrderedDict()
        layers['layer_input'] = torch.nn.Linear(num_inputs, n_neurons[0])
        for index, (n, act) in enumerate(zip(n_neurons, activations)):
            layers[f'layer_{index}'] = torch.nn.Linear(n, n)
            layers[f'activation_layer_{index}'] = get_activation(act)()
        layers['layer_output'] = torch.nn.Linear(n_neurons[-1], output_dim)
        layers['layer_unflatten'] = torch.nn.Unflatten(reshape_axis, reshape_dim)
        model = torch.nn.Sequential(layers)
        loss = torch.nn.MSELoss()
        optim = None
        if with_optim_lr is not None:                                  # G1: main_sfdqn_torch.py:52-56
            optim = torch.optim.Adam(model.parameters(), lr=with_optim_lr)
        return model, loss, optim

    return sf_model_lambda


HYPER = {"learning_rate_sf": 1e-3, "learning_rate_w": 1e-3, "learning_rate_g": 1e-3, "learning_rate_h": 1e-3,
         "weight_decay_sf": 0, "weight_decay_w": 0, "weight_decay_g": 0, "weight_decay_h": 0,
         "g_h_function_dims": 100, "beta_loss_coefficient": 1}


def linears(module):
    return [m for m in module.modules() if isinstance(m, torch.nn.Linear)]


def dump_net(out, prefix, module):
    for l, lin in enumerate(linears(module)):
        out[f'{prefix}.W{l}'] = lin.weight.detach().numpy().copy()
        out[f'{prefix}.b{l}'] = lin.bias.detach().numpy().copy()


def dump_adam(out, prefix, optim, named_params):
    """named_params: list of (name, tensor) in the order we want them keyed."""
    for name, p in named_params:
        st = optim.state.get(p, None)
        if st is None or len(st) == 0:
            continue
        out[f'{prefix}.{name}.m'] = st['exp_avg'].numpy().copy()
        out[f'{prefix}.{name}.v'] = st['exp_avg_sq'].numpy().copy()
        out[f'{prefix}.{name}.step'] = np.array(float(st['step']))


def psi_named(model):
    named = []
    for l, lin in enumerate(linears(model)):
        named += [(f'W{l}', lin.weight), (f'b{l}', lin.bias)]
    return named


def dump_transitions(out, tr, five=False):
    names = ['states', 'actions', 'phis', 'next_states', 'gammas'] if five else \
        ['states', 'actions', 'rs', 'phis', 'next_states', 'gammas']
    for n, t in zip(names, tr):
        out[f'tr.{n}'] = t.numpy().copy()


def case_g2(name, S, A, D, hidden, acts, N, B, K, policy, use_gpi, seed, target_update_ev=1000, hopper=False):
    torch.manual_seed(seed)
    sf = ref_sfdqn.DeepSF(pytorch_model_handle=make_model_lambda(hidden, acts), use_true_reward=False,
                          target_update_ev=target_update_ev, hyperparameters=dict(HYPER))
    sf.reset()
    for i in range(N):
        sf.add_training_task(FakeTask(S, A, D, i))
    out = {'meta': np.array(repr(dict(kind='g2', S=S, A=A, D=D, hidden=list(hidden), acts=list(acts), N=N, B=B, K=K,
                                       policy=policy, use_gpi=use_gpi, target_update_ev=target_update_ev)))}
    for i in range(N):
        dump_net(out, f'init.psi{i}', sf.psi[i][0][0])
        out[f'init.w{i}'] = sf.fit_w[i].weight.detach().numpy().copy()
    gen = torch.Generator().manual_seed(seed + 1)
    trs = [synthetic_transitions(B, S, A, D, gen, hopper=hopper) for _ in range(K)]
    # forward / GPI outputs on the first batch's states before any update
    with torch.no_grad():
        out['out.psi_all'] = sf.get_successors(trs[0][0]).numpy().copy()
        q, task = sf.GPI(trs[0][0], policy)
        out['out.q'] = q.numpy().copy()
        out['out.task'] = task.numpy().copy()
        q1, task1 = sf.GPI(trs[0][0][:1], policy)                       # batch-1 call: 0-dim task (sfdqn.py:239)
        out['out.q_b1'] = q1.numpy().copy()
        out['out.task_b1'] = task1.numpy().copy()
    losses = []
    for k in range(K):
        dump_transitions(out, trs[k]) if k == 0 else None
        for n_, t in zip(['states', 'actions', 'rs', 'phis', 'next_states', 'gammas'], trs[k]):
            out[f'tr{k}.{n_}'] = t.numpy().copy()
        with contextlib.redirect_stdout(io.StringIO()):
            l = sf.update_successor(trs[k], policy, use_gpi)
        losses.append([float(x) for x in l])
    out['out.losses'] = np.array(losses, dtype=np.float64)
    model, _, optim = sf.psi[policy][0]
    dump_net(out, 'post.psi', model)
    dump_net(out, 'post.tgt', sf.psi[policy][1][0])
    out['post.w'] = sf.fit_w[policy].weight.detach().numpy().copy()
    dump_adam(out, 'post.adam', optim, psi_named(model) + [('w', sf.fit_w[policy].weight)])
    out['post.updates_since_target_updated'] = np.array(sf.updates_since_target_updated)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'losses', losses[-1])


def case_g3(name, S, A, D, hidden, acts, N, B, K, policy, use_gpi, seed, gdim, beta, policies=None):
    torch.manual_seed(seed)
    hyper = dict(HYPER, g_h_function_dims=gdim, beta_loss_coefficient=beta)
    with contextlib.redirect_stdout(io.StringIO()):
        dsf = ref_tsfdqn.DeepTSF(pytorch_model_handle=make_model_lambda(hidden, acts), use_true_reward=False,
                                 target_update_ev=1000, hyperparameters=hyper)
        ag = ref_tsfdqn.TSFDQN(deep_sf=dsf, buffer_handle=lambda: ref_tsfdqn.ReplayBuffer(), gamma=0.9, T=500,
                               encoding=None, use_gpi=use_gpi, hyperparameters=hyper)
        ag.reset()
        for i in range(N):
            ag.add_training_task(FakeTask(S, A, D, i))
    policies = [policy] * K if policies is None else policies
    out = {'meta': np.array(repr(dict(kind='g3', S=S, A=A, D=D, hidden=list(hidden), acts=list(acts), N=N, B=B, K=K,
                                       policy=policy, policies=list(policies), use_gpi=use_gpi, gdim=gdim, beta=beta)))}
    for i in range(N):
        dump_net(out, f'init.psi{i}', dsf.psi[i][0][0])
        out[f'init.w{i}'] = dsf.fit_w[i].weight.detach().numpy().copy()
        out[f'init.g{i}.W'] = ag.g_functions[i].weight.detach().numpy().copy()
        out[f'init.g{i}.b'] = ag.g_functions[i].bias.detach().numpy().copy()
    out['init.h.W'] = ag.h_function.weight.detach().numpy().copy()
    out['init.h.b'] = ag.h_function.bias.detach().numpy().copy()
    gen = torch.Generator().manual_seed(seed + 1)
    trs = [synthetic_transitions(B, S, A, D, gen) for _ in range(K)]
    with torch.no_grad():
        out['out.next_psi_all'] = dsf.get_next_successors(trs[0][0]).numpy().copy()
    losses = []
    for k in range(K):
        for n_, t in zip(['states', 'actions', 'rs', 'phis', 'next_states', 'gammas'], trs[k]):
            out[f'tr{k}.{n_}'] = t.numpy().copy()
        with contextlib.redirect_stdout(io.StringIO()):
            l = ag.update_successor(trs[k], policies[k], use_gpi)
        losses.append([float(x) for x in l])
    out['out.losses'] = np.array(losses, dtype=np.float64)
    for i in sorted(set(policies)):
        model, _, optim = dsf.psi[i][0]
        dump_net(out, f'post.psi{i}', model)
        out[f'post.w{i}'] = dsf.fit_w[i].weight.detach().numpy().copy()
        out[f'post.g{i}.W'] = ag.g_functions[i].weight.detach().numpy().copy()
        out[f'post.g{i}.b'] = ag.g_functions[i].bias.detach().numpy().copy()
        dump_adam(out, f'post.adam{i}', optim, psi_named(model) + [
            ('w', dsf.fit_w[i].weight), ('gW', ag.g_functions[i].weight), ('gb', ag.g_functions[i].bias),
            ('hW', ag.h_function.weight), ('hb', ag.h_function.bias)])
    out['post.h.W'] = ag.h_function.weight.detach().numpy().copy()
    out['post.h.b'] = ag.h_function.bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'losses', losses[-1])


class FeatTask(FakeTask):
    """FakeTask with a deterministic, non-trivial feature map (the target-task loop multiplies phi into the loss)."""

    def features(self, s, a, s1):
        base = torch.arange(self.D, dtype=torch.float32)
        return torch.sin(base * 0.7 + float(torch.as_tensor(s).sum()) + 0.3 * float(a)) - 0.25 * float(torch.as_tensor(s1).sum())


TARGET_HYPER = dict(learning_rate_omega=1e-2, weight_decay_omega=0.0, learning_rate_omega_decay=1e-3, omegas_l1_coefficient=0.1)


def case_g3_target(name, S, A, D, hidden, acts, N, K, seed, gdim, beta):
    """TSFDQN target-task adaptation (tsfdqn.py:859-997): get_test_action + update_test_reward_mapper."""
    torch.manual_seed(seed)
    hyper = dict(HYPER, g_h_function_dims=gdim, beta_loss_coefficient=beta, **TARGET_HYPER)
    with contextlib.redirect_stdout(io.StringIO()):
        dsf = ref_tsfdqn.DeepTSF(pytorch_model_handle=make_model_lambda(hidden, acts), use_true_reward=False,
                                 target_update_ev=1000, hyperparameters=hyper)
        ag = ref_tsfdqn.TSFDQN(deep_sf=dsf, buffer_handle=lambda: ref_tsfdqn.ReplayBuffer(), gamma=0.9, T=500,
                               encoding=None, use_gpi=True, test_epsilon=0.0, hyperparameters=hyper)
        ag.reset()
        for i in range(N):
            ag.add_training_task(FakeTask(S, A, D, i))
    ag.total_training_steps = 1                                           # keeps the reference's debug print silent
    out = {'meta': np.array(repr(dict(kind='g3_target', S=S, A=A, D=D, hidden=list(hidden), acts=list(acts), N=N, K=K,
                                       gdim=gdim, beta=beta, gamma=0.9, **TARGET_HYPER)))}
    for i in range(N):
        dump_net(out, f'init.psi{i}', dsf.psi[i][0][0])
        out[f'init.w{i}'] = dsf.fit_w[i].weight.detach().numpy().copy()
        out[f'init.g{i}.W'] = ag.g_functions[i].weight.detach().numpy().copy()
        out[f'init.g{i}.b'] = ag.g_functions[i].bias.detach().numpy().copy()
    out['init.h.W'] = ag.h_function.weight.detach().numpy().copy()
    out['init.h.b'] = ag.h_function.bias.detach().numpy().copy()
    # what TSFDQN.train does for one test task (tsfdqn.py:797-832)
    task = FeatTask(S, A, D, 0)
    omegas = ag._init_omega(N)
    with torch.no_grad():
        omegas = omegas / torch.sum(omegas, axis=1, keepdim=True)
    omegas = omegas.clone().detach().requires_grad_(True)
    w_approx = torch.nn.Linear(D, 1, bias=False)
    with torch.no_grad():
        w_approx.weight = torch.nn.Parameter(torch.Tensor(1, D).uniform_(-0.01, 0.01))
    optim = torch.optim.Adam([
        {'params': w_approx.parameters(), 'lr': hyper['learning_rate_w'], 'weight_decay': hyper['weight_decay_w']},
        {'params': omegas, 'lr': hyper['learning_rate_omega'], 'weight_decay': hyper['weight_decay_omega']}])
    sched = torch.optim.lr_scheduler.LambdaLR(optim, [lambda e: 1 ** e, lambda e: (1 - hyper['learning_rate_omega_decay']) ** e])
    out['init.omegas'] = omegas.detach().numpy().copy()
    out['init.w_target'] = w_approx.weight.detach().numpy().copy()
    gen = torch.Generator().manual_seed(seed + 1)
    losses = []
    for k in range(K):
        s, s1 = torch.randn(1, S, generator=gen), torch.randn(1, S, generator=gen)
        r = float(torch.randn((), generator=gen))
        with torch.no_grad():
            a = ag.get_test_action(s, w_approx, omegas)
            a1 = ag.get_test_action(s1, w_approx, omegas)
        out[f'step{k}.s'], out[f'step{k}.s1'] = s.numpy().copy(), s1.numpy().copy()
        out[f'step{k}.r'], out[f'step{k}.a'], out[f'step{k}.a1'] = np.array(r), np.array(int(a)), np.array(int(a1))
        out[f'step{k}.phi'] = task.features(s, a, s1).numpy().copy()
        with contextlib.redirect_stdout(io.StringIO()):
            l = ag.update_test_reward_mapper(w_approx, omegas, optim, task, r, s, a, s1, a1)
        sched.step()
        losses.append([float(x) for x in l])
        out[f'step{k}.omegas'] = omegas.detach().numpy().copy()
        out[f'step{k}.w_target'] = w_approx.weight.detach().numpy().copy()
    out['out.losses'] = np.array(losses, dtype=np.float64)                # (loss, l2 = reward loss, l1 = psi loss)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name, 'losses', losses[-1])


def case_g1(name, S, A, D, hidden, acts, N, B, seed, lr=1e-3):
    """G1 ensemble (features/deep.py + agents/sfdqn.py:57-60): literal sequential loop AND frozen-snapshot variant."""
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        sf = RefDeepSF_G1(pytorch_model_handle=make_model_lambda(hidden, acts, with_optim_lr=lr), target_update_ev=1000,
                          use_true_reward=False, hyperparameters={'learning_rate_w': 0.5})
        sf.reset()
        for i in range(N):
            sf.add_training_task(FakeTask(S, A, D, i))
    out = {'meta': np.array(repr(dict(kind='g1', S=S, A=A, D=D, hidden=list(hidden), acts=list(acts), N=N, B=B, lr=lr)))}
    for i in range(N):
        dump_net(out, f'init.psi{i}', sf.psi[i][0][0])
        out[f'init.w{i}'] = sf.fit_w[i].numpy().copy()                    # raw tensor [D,1] (features/successor.py)
    gen = torch.Generator().manual_seed(seed + 1)
    tr = synthetic_transitions(B, S, A, D, gen, five_tuple=True)
    for n_, t in zip(['states', 'actions', 'phis', 'next_states', 'gammas'], tr):
        out[f'tr0.{n_}'] = t.numpy().copy()
    # frozen snapshot: deepcopy the reference library once per task, step copy i, collect net i from copy i
    for i in range(N):
        c = copy.deepcopy(sf)
        c.update_successor(tr, i)
        dump_net(out, f'post_frozen.psi{i}', c.psi[i][0][0])
    # literal loop
    for i in range(N):
        sf.update_successor(tr, i)
    for i in range(N):
        dump_net(out, f'post_seq.psi{i}', sf.psi[i][0][0])
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print('wrote', name)


if __name__ == '__main__':
    # Reacher shapes (reacher.cfg) with a narrow MLP to keep fixtures small, plus one full-width case
    case_g2('g2_reacher_gpi', 4, 9, 12, (64, 64), ('relu', 'relu'), N=3, B=32, K=3, policy=1, use_gpi=True, seed=1024)
    case_g2('g2_reacher_nogpi', 4, 9, 12, (64, 64), ('relu', 'relu'), N=3, B=32, K=2, policy=2, use_gpi=False, seed=1025)
    case_g2('g2_reacher_sync', 4, 9, 12, (32, 32), ('relu', 'relu'), N=2, B=16, K=3, policy=0, use_gpi=True, seed=1026,
            target_update_ev=2)
    case_g2('g2_reacher_h256', 4, 9, 12, (256, 256), ('relu', 'relu'), N=2, B=64, K=1, policy=0, use_gpi=True, seed=1027)
    case_g2('g2_cartpole_tanh', 4, 2, 20, (32, 32), ('tanh', 'tanh'), N=3, B=32, K=2, policy=0, use_gpi=False, seed=1028)
    case_g2('g2_hopper_gpi', 11, 27, 50, (32, 32), ('relu', 'relu'), N=4, B=48, K=1, policy=3, use_gpi=True, seed=1029,
            hopper=True)
    case_g3('g3_reacher_gpi', 4, 9, 12, (64, 64), ('relu', 'relu'), N=3, B=32, K=3, policy=1, use_gpi=True, seed=2024,
            gdim=100, beta=1)
    case_g3('g3_reacher_beta30', 4, 9, 12, (64, 64), ('relu', 'relu'), N=3, B=32, K=4, policy=0, use_gpi=False, seed=2025,
            gdim=100, beta=30, policies=[0, 2, 0, 2])                    # shared h stepped by two different optimizers
    case_g3_target('g3_reacher_target_adapt', 4, 9, 12, (64, 64), ('relu', 'relu'), N=3, K=5, seed=2026, gdim=100, beta=30)
    case_g1('g1_reacher_ensemble', 4, 9, 12, (64, 64), ('relu', 'relu'), N=3, B=32, seed=3024)
