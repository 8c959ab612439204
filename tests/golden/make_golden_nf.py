# -*- coding: UTF-8 -*-
"""
Generates tests/golden/g3_nf_reacher.npz by executing the UNMODIFIED reference normalising-flow TSF agent
(/root/reference/source/tsfdqn_nf.py: DeepTSF, TSFDQN, PlanarFlow) on seeded synthetic batches.  Run once in the build
container:  python tests/golden/make_golden_nf.py
"""
import contextlib
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
import tsfdqn_nf as ref                                            # noqa: E402

from tests.synthetic import synthetic_transitions                  # noqa: E402

torch.set_num_threads(1)
HYPER = {"learning_rate_sf": 1e-3, "learning_rate_w": 1e-3, "learning_rate_g": 1e-3, "learning_rate_h": 1e-3,
         "weight_decay_sf": 0, "weight_decay_w": 0, "weight_decay_g": 0, "weight_decay_h": 0}


class FakeTask:
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


def sf_lambda(hidden, acts):
    def handle(num_inputs, output_dim, reshape_dim, reshape_axis=1):
        layers = OrderedDict()
        layers['layer_input'] = torch.nn.Linear(num_inputs, hidden[0])
        for k, (n, a) in enumerate(zip(hidden, acts)):
            layers[f'layer_{k}'] = torch.nn.Linear(n, n)
            layers[f'activation_layer_{k}'] = get_activation(a)()
        layers['layer_output'] = torch.nn.Linear(hidden[-1], output_dim)
        layers['layer_unflatten'] = torch.nn.Unflatten(reshape_axis, reshape_dim)
        return torch.nn.Sequential(layers), torch.nn.MSELoss(), None
    return handle


def lin(m):
    return [x for x in m.modules() if isinstance(x, torch.nn.Linear)]


def dump_g(out, prefix, g):
    mods = list(g)
    out[f'{prefix}.n_flows'] = len(mods) - 1
    for k, f in enumerate(mods[:-1]):
        out[f'{prefix}.f{k}.weight'] = f.weight.detach().numpy().copy()
        out[f'{prefix}.f{k}.bias'] = f.bias.detach().numpy().copy()
        out[f'{prefix}.f{k}.scale'] = f.scale.detach().numpy().copy()
    out[f'{prefix}.W'], out[f'{prefix}.b'] = mods[-1].weight.detach().numpy().copy(), mods[-1].bias.detach().numpy().copy()


def case(name, S, A, D, hidden, acts, N, B, K, policy, use_gpi, seed, gdim, beta, n_flows):
    torch.manual_seed(seed)
    hyper = dict(HYPER, g_h_function_dims=gdim, beta_loss_coefficient=beta, n_coupling_layers=n_flows)
    with contextlib.redirect_stdout(io.StringIO()):
        dsf = ref.DeepTSF(pytorch_model_handle=sf_lambda(hidden, acts), use_true_reward=False, target_update_ev=1000,
                          hyperparameters=hyper)
        ag = ref.TSFDQN(deep_sf=dsf, buffer_handle=lambda: ref.ReplayBuffer(), gamma=0.9, T=500, encoding=None, use_gpi=use_gpi,
                        hyperparameters=hyper)
        ag.reset()
        for i in range(N):
            ag.add_training_task(FakeTask(S, A, D, i))
    out = dict(S=S, A=A, D=D, hidden=np.asarray(hidden), acts=np.asarray(acts), N=N, B=B, K=K, policy=policy, use_gpi=int(use_gpi),
               gdim=gdim, beta=beta, n_flows=n_flows)
    for i in range(N):
        for l, m in enumerate(lin(dsf.psi[i][0][0])):
            out[f'init.psi{i}.W{l}'], out[f'init.psi{i}.b{l}'] = m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()
        out[f'init.w{i}'] = dsf.fit_w[i].weight.detach().numpy().copy()
        dump_g(out, f'init.g{i}', ag.g_functions[i])
    out['init.h.W'], out['init.h.b'] = ag.h_function.weight.detach().numpy().copy(), ag.h_function.bias.detach().numpy().copy()
    n_trainable = sum(1 for _ in ag.g_functions[policy].parameters())
    out['g_trainable_tensors'] = n_trainable                         # 3 * n_flows + 2 on CPU (see oracle note on `.to(device)`)
    gen = torch.Generator().manual_seed(seed + 1)
    losses = []
    for k in range(K):
        tr = synthetic_transitions(B, S, A, D, gen)
        for n_, t_ in zip(('states', 'actions', 'rs', 'phis', 'next_states', 'gammas'), tr):
            out[f'tr{k}.{n_}'] = t_.numpy().copy()
        with contextlib.redirect_stdout(io.StringIO()):
            l = ag.update_successor(tr, policy, use_gpi)
        losses.append([float(x) for x in l])
    out['out.losses'] = np.asarray(losses)
    for l, m in enumerate(lin(dsf.psi[policy][0][0])):
        out[f'post.psi.W{l}'], out[f'post.psi.b{l}'] = m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()
    out['post.w'] = dsf.fit_w[policy].weight.detach().numpy().copy()
    dump_g(out, 'post.g', ag.g_functions[policy])
    out['post.h.W'], out['post.h.b'] = ag.h_function.weight.detach().numpy().copy(), ag.h_function.bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'g tensors trained:', n_trainable, 'losses', losses)


if __name__ == '__main__':
    case('g3_nf_reacher', S=4, A=9, D=12, hidden=(64, 64), acts=('relu', 'relu'), N=3, B=32, K=4, policy=1, use_gpi=True, seed=93,
         gdim=100, beta=1, n_flows=5)
