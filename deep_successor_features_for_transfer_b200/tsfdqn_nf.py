# -*- coding: UTF-8 -*-
"""
Normalising-flow TSF agent: host mirror of the reference's tsfdqn_nf.py.  That file is tsfdqn.py with ONE difference -- the
per-task g function is a chain of planar flows followed by the Linear(S, G) (tsfdqn_nf.py:331-358, 569-571, 763-765; the number
of flows is the hyper-parameter `n_coupling_layers`) -- so everything else is inherited from tsfdqn.py here as well: same class
names, constructor arguments and methods (DeepTSF, TSFDQN, ReplayBuffer, PlanarFlow.build_planar_flow).

On the device the flows' parameters live behind the Linear in each policy's packed g row (W | b | K x (weight | bias | scale),
library.PackedSFLibrary) and the train step is the same fused launch chain: the TD kernel (csrc/td.cu, sfgpi_td_args.n_flows)
runs both state rows of every transition through the flows before the affine map h(g(.)) = M z + c and sweeps back through them
for the flows' gradients; Adam steps the whole g row as one segment.  The modules returned through `g_functions[i]` hold
views of that row.
"""
import torch

from .tsfdqn import DeepTSF, ReplayBuffer, TSFDQN as _TSFDQN          # noqa: F401  (re-exported: the reference module's surface)


class PlanarFlow(torch.nn.Module):
    """z -> z + scale * tanh(z . weight + bias)   [tsfdqn_nf.py:331-349].  Parameters are registered (the reference's
    `Parameter(...).to(device)` keeps them registered on CPU, where its fixture was generated, so the optimizer trains them)."""

    def __init__(self, dim, device=None):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.empty(1, dim, device=device))
        self.bias = torch.nn.Parameter(torch.empty(1, device=device))
        self.scale = torch.nn.Parameter(torch.empty(1, dim, device=device))
        self.tanh = torch.nn.Tanh()
        self.reset_parameters()

    def reset_parameters(self):                               # same draws, same order as the reference (CPU generator)
        for t in (self.weight, self.scale, self.bias):
            t.data.copy_(torch.empty(t.shape).uniform_(-0.01, 0.01))

    def forward(self, z):
        return z + self.scale * self.tanh(torch.nn.functional.linear(z, self.weight, self.bias))

    @classmethod
    def build_planar_flow(cls, input_dim, output_dim, n_affine_flows, device=None):
        flows = [cls(input_dim, device) for _ in range(n_affine_flows)]
        flows.append(torch.nn.Linear(input_dim, output_dim, bias=True, device=device))
        return torch.nn.Sequential(*flows)


class TSFDQN(_TSFDQN):
    """TSFDQN whose g_i are planar-flow chains  [tsfdqn_nf.py:569-571, 763-765]."""

    def _init_g_function(self, states_dim, output_dim, n_coupling_layers=None):
        if n_coupling_layers is None:
            n_coupling_layers = self.hyperparameters.get('n_coupling_layers', 1)
        return PlanarFlow.build_planar_flow(states_dim, output_dim, n_coupling_layers, self.device)
