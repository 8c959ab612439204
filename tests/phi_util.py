"""Deterministic stand-in task for the phi pre-training tests (the reference's tasks need pybullet / gym)."""
import math

import torch


class FakePhiTask:
    """Task protocol consumed by SFDQN.pre_train (sfdqn_phi.py:800-873): seeded smooth dynamics, rewards in [-1, 1]."""

    def __init__(self, S, A, D, index, horizon=17):
        self.S, self.A, self.D, self.index, self.horizon = S, A, D, index, horizon
        g = torch.Generator().manual_seed(1000 + index)
        self.M = torch.randn(A, S, S, generator=g) * 0.6
        self.c = torch.randn(A, S, generator=g) * 0.3
        self.u = torch.randn(S, generator=g)
        self.s, self.t = None, 0

    def action_count(self):
        return self.A

    def action_dim(self):
        return 1

    def feature_dim(self):
        return self.D

    def encode_dim(self):
        return self.S

    def initialize(self):
        self.t = 0
        self.s = torch.linspace(-1.0, 1.0, self.S) * (0.5 + 0.1 * self.index)
        return self.s

    def encode(self, s):
        return s.reshape(1, -1).float()

    def transition(self, a):
        self.s = torch.tanh(self.M[a] @ self.s + self.c[a])
        self.t += 1
        r = math.tanh(float(self.u @ self.s) + 0.1 * a)
        return self.s, r, self.t >= self.horizon
