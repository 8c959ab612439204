"""
A deterministic stand-in environment implementing the reference's task protocol (tasks/task.py: initialize, transition,
features, encode, action_count, feature_dim, encode_dim, get_w) so that the AGENT-level loops -- SFDQN/TSFDQN
next_sample -> train_agent -> replay -> update_successor (sfdqn.py:462-484, 550-627; tsfdqn.py:435-497, 566-586) -- can be
run identically by the unmodified reference (tests/golden/make_golden_agent.py) and by the CUDA path (tests/test_gpu_agent.py).
The real tasks (pybullet Reacher, gym Hopper / CartPole) are host-side rollouts and out of scope; only their call protocol and
tensor types matter here: states are float32 tensors [1, S], actions 0-dim int64 tensors, rewards float64 0-dim tensors
("rewards are double", sfdqn.py:86), features float32 [D].
"""
import numpy as np
import torch


class ToyTask:
    def __init__(self, S, A, D, index, seed=0, episode_len=17):
        self.S, self.A, self.D, self.index = S, A, D, index
        self.rng = np.random.RandomState(1000 * seed + index)
        self.dirs = self.rng.uniform(-0.5, 0.5, size=(A, S)).astype(np.float32)
        self.centres = self.rng.uniform(-1.0, 1.0, size=(D, S)).astype(np.float32)
        self.w = np.zeros((D, 1), dtype=np.float32)
        self.w[index % D, 0] = 1.0
        self.episode_len = episode_len
        self.s = None
        self.t = 0
        self.actions_taken = []

    # ---- shape protocol ----
    def action_count(self):
        return self.A

    def feature_dim(self):
        return self.D

    def encode_dim(self):
        return self.S

    def get_w(self):
        return torch.from_numpy(self.w.copy())

    def encode(self, s):
        return s

    # ---- dynamics ----
    def initialize(self):
        self.s = self.rng.uniform(-1.0, 1.0, size=(1, self.S)).astype(np.float32)
        self.t = 0
        return torch.from_numpy(self.s.copy())

    def features(self, s, a, s1):
        """phi(s, a, s') = 1 - 4 * distance of s' to D fixed centres (the Reacher form, tasks/reacher.py:79), float32 [D]."""
        s1 = torch.as_tensor(s1).detach().cpu().float().reshape(1, self.S)
        d = torch.linalg.norm(s1 - torch.from_numpy(self.centres), dim=1)
        return (1.0 - 4.0 * d).float()

    def transition(self, a):
        a = int(a)
        self.actions_taken.append(a)
        noise = self.rng.normal(0.0, 0.05, size=(1, self.S)).astype(np.float32)
        s1 = np.clip(0.9 * self.s + self.dirs[a][None, :] + noise, -2.0, 2.0).astype(np.float32)
        phi = self.features(None, a, torch.from_numpy(s1))
        r = (phi.double() @ torch.from_numpy(self.w).double()).reshape(())          # double, like the reference's tasks
        self.s = s1
        self.t += 1
        terminal = self.t >= self.episode_len
        return torch.from_numpy(s1.copy()), r, terminal
