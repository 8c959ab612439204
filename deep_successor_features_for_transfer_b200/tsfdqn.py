# -*- coding: UTF-8 -*-
"""
Drop-in for the reference's single-file G3 API (`source/tsfdqn.py`): ReplayBuffer, DeepTSF, TSFDQN (Transformed Successor
Features).  The TSF train step -- phi~ = phi * (h(g_i(s)) + h(g_i(s'))), target phi~ + gamma * psi^-(s', a*), loss
l1 + beta * l2, Adam over (psi_i, w_i, g_i, h) -- is one fused kernel sequence in libsfgpi.so (tsfdqn.py:588-709).
"""
import random

import numpy as np
import torch

from .sfdqn import DeepSF, DeviceReplayBuffer, ReplayBuffer, SFDQN, _device  # noqa: F401  (re-exported like the reference file)


class DeepTSF(DeepSF):
    """SF library of the TSF agent  [tsfdqn.py:93-324]: adds target-net forwards and the g/h optimizer groups."""

    def __init__(self, pytorch_model_handle, use_true_reward=False, target_update_ev=1000, **kwargs):
        super().__init__(pytorch_model_handle, use_true_reward, target_update_ev, **kwargs)
        self._tsf_dim = self.hyperparameters.get('g_h_function_dims')

    def add_training_task(self, task, source=None, g_function_model=None, h_function_model=None):
        if g_function_model is None or h_function_model is None:
            raise Exception('DeepTSF.add_training_task needs the g and h functions (tsfdqn.py:137)')
        if self._tsf_dim is None:
            self._tsf_dim = g_function_model.out_features
            self._library = self._new_library()
        true_w = task.get_w()
        n_features = task.feature_dim()
        w_approx = torch.nn.Linear(n_features, 1, bias=False, device=self.device)
        with torch.no_grad():
            w_approx.weight.uniform_(-0.01, 0.01)
        self.true_w.append(true_w)
        self.fit_w.append(w_approx)
        self.psi.append(self.build_successor(task, source, w_approx, g_function_model, h_function_model))
        self.n_tasks = len(self.psi)
        for i in range(len(self.gpi_counters)):
            self.gpi_counters[i] = np.append(self.gpi_counters[i], 0)
        self.gpi_counters.append(np.zeros((self.n_tasks,), dtype=int))

    def get_next_successor(self, state, policy_index):
        return self._library.forward_psi(state, policy_index, 1, target=True)[:, 0]

    def get_next_successors(self, state):
        return self._library.forward_psi(state, 0, self.n_tasks, target=True)

    def update_successor(self, transitions, policy_index, use_gpi=True):
        raise Exception('This function should not be called')      # features/deep_sequential_tsf.py:184-185


class TSFDQN(SFDQN):
    """TSFDQN agent  [tsfdqn.py:329-1011]."""

    def __init__(self, deep_sf, buffer_handle, gamma, T, encoding, epsilon=0.1, epsilon_decay=1., epsilon_min=0.,
                 print_ev=1000, save_ev=100, use_gpi=True, test_epsilon=0.03, **kwargs):
        super().__init__(deep_sf, buffer_handle, gamma, T, encoding, epsilon, epsilon_decay, epsilon_min, print_ev, save_ev,
                         use_gpi, test_epsilon, **kwargs)
        self.omegas_per_source_task = []
        self.omegas = []
        self.g_functions = []
        self.h_function = None

    def reset(self):
        super().reset()
        self.g_functions = []
        self.h_function = None

    # ---- g / h / omega (tsfdqn.py:537-564) --------------------------------------------------------------------------
    def _init_g_function(self, states_dim, output_dim):
        return torch.nn.Linear(states_dim, output_dim, bias=True, device=self.device)

    def _init_h_function(self, input_dim, features_dim):
        return torch.nn.Linear(input_dim, features_dim, bias=True, device=self.device)

    def _init_omega(self, num_source_tasks):
        return torch.Tensor(1, num_source_tasks, 1, 1).uniform_(0, 1).to(self.device).requires_grad_(True)

    def add_training_task(self, task):
        self.tasks.append(task)
        self.n_tasks = len(self.tasks)
        self.phis.append(task.features)
        if self.n_tasks == 1:
            self.n_actions = task.action_count()
            self.n_features = task.feature_dim()
            if self.encoding == 'task':
                self.encoding = task.encode
        self.buffers.append(self.buffer_handle())
        g_h_function_dims = self.hyperparameters.get('g_h_function_dims')
        g_function = self._init_g_function(task.encode_dim(), g_h_function_dims)
        self.g_functions.append(g_function)
        if self.h_function is None:
            self.h_function = self._init_h_function(g_h_function_dims, task.feature_dim())
        self.sf.add_training_task(task, None, g_function, self.h_function)

    def set_active_training_task(self, index):
        super().set_active_training_task(index)
        self.active_g_function = self.g_functions[index]

    def get_Q_values(self, s, s_enc):
        with torch.no_grad():
            q, c = self.sf.GPI(s_enc, self.task_index, update_counters=self.use_gpi)
            if not self.use_gpi:
                c = self.task_index
            self.c = c
            return q[:, c, :]

    def _greedy_action(self):
        q = self.get_Q_values(self.s, self.s_enc).flatten()
        assert q.size()[0] == self.n_actions
        return torch.argmax(q)

    # ---- the TSF train step -------------------------------------------------------------------------------------------
    def train_agent(self, s, s_enc, a, r, s1, s1_enc, gamma):
        phi = self.phi(s, a, s1)
        self.buffer.append(s_enc, a, r, phi, s1_enc, gamma)
        transitions = self.buffer.replay()
        losses = self.update_successor(transitions, self.task_index, self.use_gpi)
        if isinstance(losses, tuple) and self.logger is not None:
            total_loss, psi_loss, phi_loss = losses
            self.logger.log_losses(total_loss.item(), psi_loss.item(), phi_loss.item(),
                                   [self.hyperparameters['beta_loss_coefficient']], self.total_training_steps)

    def update_successor(self, transitions, policy_index, use_gpi=True):
        if transitions is None:
            return
        if self.h_function is None:
            raise Exception('Affine Function (h) is not initialized')
        beta = self.hyperparameters['beta_loss_coefficient']
        losses = self.sf._library.train_step(transitions, policy_index, use_gpi=use_gpi, variant=2, beta=beta)
        self.sf._after_update(policy_index)
        return losses[0, 0], losses[0, 1], losses[0, 2]

    def update_successor_all(self, transitions, use_gpi=True):
        """Ensemble extension (BASELINE config 4-ii): all policies stepped on one batch, frozen-snapshot semantics."""
        if transitions is None:
            return
        beta = self.hyperparameters['beta_loss_coefficient']
        losses = self.sf._library.train_step(transitions, 'all', use_gpi=use_gpi, variant=2, beta=beta)
        for i in range(self.sf.n_tasks):
            self.sf._after_update(i)
        return losses
