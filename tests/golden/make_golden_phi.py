# -*- coding: UTF-8 -*-
"""
Generates tests/golden/phi_pretrain_cartpole.npz by executing the UNMODIFIED reference `SFDQN.pre_train`
(/root/reference/source/sfdqn_phi.py:800-873, with its own PhiFunction and ReplayBuffer) on the deterministic FakePhiTask
of tests/phi_util.py.  Run once in the build container:  python tests/golden/make_golden_phi.py

Stored: the initial weights the reference draws under the seed (PhiFunction layers, the per-task reward heads), the list of
losses pre_train returns (one per update), the final PhiFunction weights and its output on a probe batch.
"""
import contextlib
import io
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference/source")
from utils.torch import set_torch_device                          # noqa: E402
from utils.logger import set_logger_level                         # noqa: E402

with contextlib.redirect_stdout(io.StringIO()):
    set_torch_device(use_gpu=False)
    set_logger_level(use_logger=False)
import sfdqn_phi as ref                                            # noqa: E402  (its main block is guarded)

from tests.phi_util import FakePhiTask                             # noqa: E402

ref.device = torch.device('cpu')                                   # the module global its main block would set (:922)
torch.set_num_threads(1)


def seed_all(seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)


def case(name, S, A, D, n_tasks, n_samples, n_cycles, seed):
    out = dict(S=S, A=A, D=D, n_tasks=n_tasks, n_samples=n_samples, n_cycles=n_cycles, seed=seed)
    # the draws pre_train makes under this seed, in its order: PhiFunction, then one head per task (:808-826)
    seed_all(seed)
    phi0 = ref.PhiFunction(S, 1, D)
    lin = [m for m in phi0._model if isinstance(m, torch.nn.Linear)]
    for l, m in enumerate(lin):
        out[f'init.phi.W{l}'] = m.weight.detach().numpy().copy()
        out[f'init.phi.b{l}'] = m.bias.detach().numpy().copy()
    for i in range(n_tasks):
        torch.nn.Linear(D, 1, bias=False)
        out[f'init.w{i}'] = torch.Tensor(1, D).uniform_(-0.01, 0.01).numpy().copy()
    # the real thing
    seed_all(seed)
    tasks = [FakePhiTask(S, A, D, i) for i in range(n_tasks)]
    holder = types.SimpleNamespace()
    losses = ref.SFDQN.pre_train(holder, tasks, n_samples, n_cycles)
    model = holder.learnt_phi
    for l, m in enumerate([m for m in model._model if isinstance(m, torch.nn.Linear)]):
        out[f'out.phi.W{l}'] = m.weight.detach().numpy().copy()
        out[f'out.phi.b{l}'] = m.bias.detach().numpy().copy()
    out['out.losses'] = np.asarray(losses, dtype=np.float64)
    g = torch.Generator().manual_seed(seed + 1)
    ps, pa, ps1 = torch.randn(16, S, generator=g), torch.randint(0, A, (16,), generator=g), torch.randn(16, S, generator=g)
    out['probe.s'], out['probe.a'], out['probe.s1'] = ps.numpy(), pa.numpy(), ps1.numpy()
    with torch.no_grad():
        out['probe.phi'] = model(ps, pa, ps1).numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, 'updates:', len(losses), 'first/last loss:', losses[0], losses[-1])


if __name__ == '__main__':
    # CartPole shapes of cartpole_phi.cfg (S = 4, 2 actions, 20 features): 9 -> 128 -> 256 -> 20
    case('phi_pretrain_cartpole', S=4, A=2, D=20, n_tasks=2, n_samples=45, n_cycles=2, seed=77)
