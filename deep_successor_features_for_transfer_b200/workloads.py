# -*- coding: UTF-8 -*-
"""
Synthetic workloads of the reference's shapes (SURVEY section 8d "Synthetic inputs"): the pieces the
benchmark, the measurement scripts and the GPU tests share to stand up a library without the reference's environments
(pybullet / gym rollouts are host-side and out of scope) -- a shapes-only task, the mains' MLP factory and replay batches
with the reference's value ranges.  Nothing here touches the CPU oracle.
"""
from collections import OrderedDict

import torch

ACTS = {'relu': torch.nn.ReLU, 'tanh': torch.nn.Tanh}

# the [SFDQN] hyper-parameters of configs/reacher.cfg:22-49 that the hot path reads
HYPER = {"learning_rate_sf": 1e-3, "learning_rate_w": 1e-3, "learning_rate_g": 1e-3, "learning_rate_h": 1e-3,
         "weight_decay_sf": 0, "weight_decay_w": 0, "weight_decay_g": 0, "weight_decay_h": 0,
         "g_h_function_dims": 100, "beta_loss_coefficient": 1}

ENVS = {
    'reacher': dict(S=4, A=9, D=12),        # tasks/reacher.py:22-26, 82-83
    'hopper': dict(S=11, A=27, D=50),       # hopper_phi.cfg:52, tasks/hopper_phi.py:64
    'cartpole': dict(S=4, A=2, D=20),       # cartpole_phi.cfg:52
}


class ShapeTask:
    """Task protocol consumed by the SF library (tasks/task.py): shapes only."""

    def __init__(self, S, A, D, index=0):
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


def model_lambda(hidden, acts):
    """Same shape contract as the mains' sf_model_lambda (main_tsfdqn_sequential_torch.py:44-75)."""

    def handle(num_inputs, output_dim, reshape_dim, reshape_axis=1):
        layers = OrderedDict()
        layers['layer_input'] = torch.nn.Linear(num_inputs, hidden[0])
        for k, (n, a) in enumerate(zip(hidden, acts)):
            layers[f'layer_{k}'] = torch.nn.Linear(n, n)
            layers[f'activation_layer_{k}'] = ACTS[a]()
        layers['layer_output'] = torch.nn.Linear(hidden[-1], output_dim)
        layers['layer_unflatten'] = torch.nn.Unflatten(reshape_axis, reshape_dim)
        return torch.nn.Sequential(layers), torch.nn.MSELoss(), None

    return handle


def synthetic_transitions(B, S, A, D, gen, hopper=False, five_tuple=False):
    states = torch.randn(B, S, generator=gen)
    next_states = torch.randn(B, S, generator=gen)
    if hopper:                                                            # tasks/hopper_phi.py:59
        states, next_states = torch.sigmoid(states), torch.sigmoid(next_states)
    actions = torch.randint(0, A, (B,), generator=gen, dtype=torch.int64)
    phis = torch.rand(B, D, generator=gen) * 2.5 - 1.5                   # U(-1.5, 1), tsfdqn.py:541
    w_true = torch.zeros(D, 1)
    w_true[0, 0] = 1.0                                                    # one-hot, tasks/reacher.py:85-88
    rs = phis @ w_true
    gammas = torch.full((B,), 0.9)
    gammas[torch.rand(B, generator=gen) < 0.01] = 0.0                    # 1 % terminals
    if five_tuple:
        return states, actions, phis, next_states, gammas
    return states, actions, rs, phis, next_states, gammas


def seeded_policy_init(seed):
    """
    Context manager: torch's global RNG seeded for the construction of ONE policy (its psi nets, w, g) so that any rank of a
    policy-sharded run -- and the unsharded check on rank 0 -- builds bit-identical weights for global policy index k from
    seed + k alone.
    """
    return torch.random.fork_rng(devices=[]) if seed is None else _Seeded(seed)


class _Seeded:
    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        self.state = torch.random.get_rng_state()
        torch.manual_seed(self.seed)

    def __exit__(self, *exc):
        torch.random.set_rng_state(self.state)


def build_tsf_agent(env, n_policies, hidden=(256, 256), acts=('relu', 'relu'), gdim=100, beta=1, precision='fp32', use_gpi=True,
                    seed=None, first_policy=0, target_update_ev=1000):
    """
    A TSFDQN agent + DeepTSF library with `n_policies` source tasks of `env`'s shapes (the object graph the reference's
    main_tsfdqn_sequential_torch.py:85-92 builds), weights from nn.Linear's default init.  seed: per-policy seeding (policy with
    global index first_policy + i is drawn from seed + first_policy + i; the shared h from seed - 1).
    """
    from .tsfdqn import DeepTSF, TSFDQN, ReplayBuffer
    shp = ENVS[env] if isinstance(env, str) else env
    hyper = dict(HYPER, g_h_function_dims=gdim, beta_loss_coefficient=beta, precision=precision)
    dsf = DeepTSF(pytorch_model_handle=model_lambda(list(hidden), list(acts)), use_true_reward=False,
                  target_update_ev=target_update_ev, hyperparameters=hyper)
    ag = TSFDQN(deep_sf=dsf, buffer_handle=lambda: ReplayBuffer(), gamma=0.9, T=500, encoding=None, use_gpi=use_gpi,
                hyperparameters=hyper)
    ag.reset()
    if seed is not None:
        with _Seeded(seed - 1):
            ag.h_function = ag._init_h_function(gdim, shp['D'])
    for i in range(n_policies):
        k = first_policy + i
        if seed is None:
            ag.add_training_task(ShapeTask(shp['S'], shp['A'], shp['D'], k))
        else:
            with _Seeded(seed + k):
                ag.add_training_task(ShapeTask(shp['S'], shp['A'], shp['D'], k))
    return dsf, ag


def build_sf_library(env, n_policies, hidden=(256, 256), acts=('relu', 'relu'), precision='fp32', seed=None, first_policy=0,
                     target_update_ev=1000):
    """A DeepSF (G2) library with `n_policies` source tasks (sfdqn.py:180-288), same seeding convention as build_tsf_agent."""
    from .sfdqn import DeepSF
    shp = ENVS[env] if isinstance(env, str) else env
    sf = DeepSF(pytorch_model_handle=model_lambda(list(hidden), list(acts)), target_update_ev=target_update_ev,
                hyperparameters=dict(HYPER, precision=precision))
    sf.reset()
    for i in range(n_policies):
        k = first_policy + i
        if seed is None:
            sf.add_training_task(ShapeTask(shp['S'], shp['A'], shp['D'], k))
        else:
            with _Seeded(seed + k):
                sf.add_training_task(ShapeTask(shp['S'], shp['A'], shp['D'], k))
    return sf
