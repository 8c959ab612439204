# -*- coding: UTF-8 -*-
"""
Generates tests/golden/g4_joint_psi_phi.npz by executing the UNMODIFIED reference G4 step -- features/deep_phi.py
DeepSF_PHI.update_successor with the phi model of main_sfdqn_phi_torch.py:52-73 and the loss coefficient of
agents/sfdqn_phi.py:152-165 -- on seeded synthetic batches.  Run once in the build container:
    python tests/golden/make_golden_g4.py
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
from features.deep_phi import DeepSF_PHI                          # noqa: E402

from tests.synthetic import synthetic_transitions                  # noqa: E402

torch.set_num_threads(1)
torch.autograd.set_detect_anomaly(False)


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
        return torch.zeros(self.D, 1)


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


def phi_lambda(s_enc_dim, action_dim, feature_dim):                   # main_sfdqn_phi_torch.py:52-73
    n = s_enc_dim * 2 + action_dim
    model = torch.nn.Sequential(torch.nn.Linear(n, 2 * n), torch.nn.ReLU(), torch.nn.Linear(2 * n, 2 * n), torch.nn.ReLU(),
                                torch.nn.Linear(2 * n, 2 * n), torch.nn.ReLU(), torch.nn.Linear(2 * n, 2 * n), torch.nn.ReLU(),
                                torch.nn.Linear(2 * n, feature_dim))
    return model, torch.nn.MSELoss(), None


def lin(m):
    return [x for x in m.modules() if isinstance(x, torch.nn.Linear)]


def case(name, S, A, D, hidden, acts, N, B, K, policy, use_gpi, seed):
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        sf = DeepSF_PHI(pytorch_model_handle=sf_lambda(hidden, acts), use_true_reward=False, target_update_ev=1000,
                        hyperparameters={'learning_rate_w': 1e-3})
        sf.reset()
        for i in range(N):
            sf.add_training_task(FakeTask(S, A, D, i))
    phi_tuple = phi_lambda(S, 1, D)
    phis_model = (phi_tuple, phi_tuple)
    coef = torch.ones(1, requires_grad=True)                           # agents/sfdqn_phi.py:163-165
    out = dict(S=S, A=A, D=D, hidden=np.asarray(hidden), acts=np.asarray(acts), N=N, B=B, K=K, policy=policy, use_gpi=int(use_gpi))
    for i in range(N):
        for l, m in enumerate(lin(sf.psi[i][0][0])):
            out[f'init.psi{i}.W{l}'], out[f'init.psi{i}.b{l}'] = m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()
        out[f'init.w{i}.W'], out[f'init.w{i}.b'] = sf.fit_w[i].weight.detach().numpy().copy(), sf.fit_w[i].bias.detach().numpy().copy()
    for l, m in enumerate(lin(phi_tuple[0])):
        out[f'init.phi.W{l}'], out[f'init.phi.b{l}'] = m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()
    gen = torch.Generator().manual_seed(seed + 1)
    losses = []
    for k in range(K):
        tr = list(synthetic_transitions(B, S, A, D, gen))
        tr[1] = tr[1].float()                                          # actions are concatenated into phi's input (:110)
        for n_, t_ in zip(('states', 'actions', 'rs', 'phis', 'next_states', 'gammas'), tr):
            out[f'tr{k}.{n_}'] = t_.numpy().copy()
        tr_ref = (tr[0], tr[1].long(), tr[2], tr[3], tr[4], tr[5])
        with contextlib.redirect_stdout(io.StringIO()):
            loss, psi_l, phi_l, c = sf.update_successor(tr_ref, phis_model, policy, coef, use_gpi)
        losses.append([float(loss), float(psi_l), float(phi_l), float(c)])
    out['out.losses'] = np.asarray(losses)
    for l, m in enumerate(lin(sf.psi[policy][0][0])):
        out[f'post.psi.W{l}'], out[f'post.psi.b{l}'] = m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()
    for l, m in enumerate(lin(phi_tuple[0])):
        out[f'post.phi.W{l}'], out[f'post.phi.b{l}'] = m.weight.detach().numpy().copy(), m.bias.detach().numpy().copy()
    out['post.w.W'], out['post.w.b'] = sf.fit_w[policy].weight.detach().numpy().copy(), sf.fit_w[policy].bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, losses)


if __name__ == '__main__':
    case('g4_joint_psi_phi', S=4, A=2, D=20, hidden=(64, 64), acts=('relu', 'relu'), N=3, B=32, K=4, policy=1, use_gpi=True, seed=91)
