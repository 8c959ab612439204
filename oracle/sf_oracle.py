# -*- coding: UTF-8 -*-
"""
CPU ORACLE for the SF/GPI hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
module, and only as the checker / the CPU arm -- never as the thing shipped.  The product
(`deep_successor_features_for_transfer_b200`) never imports it.

It restates, function by function, the reference algorithm of okgarces/deep-successor-features-for-transfer
with plain fp32 torch CPU tensors (no nn.Module, no torch.optim): every function cites the reference file:line it
follows.  The arithmetic the reference delegates to a third-party dependency -- PyTorch (`aten::addmm`, `mse_loss`,
autograd, `torch.optim.Adam`; NOT vendored under the reference, no version pin anywhere in it; the installed
torch 2.11.0 is the de-facto pin) -- is restated here explicitly (Adam: torch/optim/adam.py::_single_tensor_adam).

Pinning: the reference ships no tests / golden vectors, so this oracle is pinned against
outputs of the reference itself, executed in the build container by `tests/golden/make_golden.py` (imports
/root/reference/source unmodified) and committed under `tests/golden/*.npz`; `tests/test_oracle_golden.py`
checks oracle == reference on those fixtures.
"""
import copy
import math

import torch

ADAM_BETA1, ADAM_BETA2, ADAM_EPS = 0.9, 0.999, 1e-8   # torch.optim.Adam defaults, used at sfdqn.py:286 / tsfdqn.py:270


# ----------------------------------------------------------------------------------------------------------------
# A1: psi network (main_tsfdqn_sequential_torch.py:44-75): Linear(S,H0) [no act] -> [Linear(H,H), act]*L -> Linear(H,A*D)
# ----------------------------------------------------------------------------------------------------------------
def act_fn(name):
    if name == 'relu':
        return torch.relu
    if name == 'tanh':
        return torch.tanh
    if name in (None, 'none'):
        return lambda t: t
    raise Exception('Activation name not supported')      # utils/torch.py:24-27


def mlp_forward(layers, acts, x):
    """layers: list of (W[out,in], b[out]); acts: activation name after each layer ('none' | 'relu' | 'tanh')."""
    h = x
    for (W, b), a in zip(layers, acts):
        h = act_fn(a)(torch.addmm(b, h, W.t()))
    return h


# ---- reduced-precision emulation of the tensor-core modes (checker for the bf16 / tf32 kernels, not a reference function) ----
# The tcgen05 kernels round every GEMM OPERAND (states, weights, stored activations, dZ) to bf16 (or tf32) and accumulate in
# fp32.  mlp_forward_emul restates exactly that on the CPU so that the tensor-core modes can be checked at a tight bound
# against an oracle with the SAME rounding points (what remains is summation order and rounding-boundary flips), in addition to
# the stated mode tolerance against the fp32 oracle.
def _round_bf16(t):
    return t.bfloat16().float()


def _round_tf32(t):
    """cvt.rna.tf32.f32: round to nearest (ties away from zero) onto 10 explicit mantissa bits."""
    bits = t.contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


ROUNDERS = {'bf16': _round_bf16, 'tf32': _round_tf32}


class _RoundOperand(torch.autograd.Function):
    """forward: the value the tensor core consumes; backward: identity (the fp32 master copy receives the gradient)."""

    @staticmethod
    def forward(ctx, t, mode):
        return ROUNDERS[mode](t)

    @staticmethod
    def backward(ctx, g):
        return g, None


class _RoundGrad(torch.autograd.Function):
    """forward: identity; backward: rounds the gradient (dZ is an MMA operand of both the dgrad chain and wgrad)."""

    @staticmethod
    def forward(ctx, t, mode):
        ctx.mode = mode
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return ROUNDERS[ctx.mode](g), None


def mlp_forward_emul(layers, acts, x, mode, upto=None):
    """mlp_forward with the tensor-core kernels' rounding points; upto = L-1 stops before the output layer (returns its input)."""
    h = _RoundOperand.apply(x, mode)
    for l, ((W, b), a) in enumerate(zip(layers, acts)):
        if upto is not None and l == upto:
            return h
        z = _RoundGrad.apply(torch.addmm(b, h, _RoundOperand.apply(W, mode).t()), mode)
        h = act_fn(a)(z)
        if l + 1 < len(layers):
            h = _RoundOperand.apply(h, mode)
    return h


def make_acts(n_hidden_layers, activations):
    """Activation pattern produced by sf_model_lambda: none after layer_input, act_k after layer_k, none after output."""
    return ['none'] + list(activations)[:n_hidden_layers] + ['none']


def init_linear(out_f, in_f, gen):
    """nn.Linear default init (kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(in)) for W and b)."""
    bound = 1.0 / math.sqrt(in_f)
    W = (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen) * 2 - 1) * bound
    return W, b


def g_apply(x, g):
    """
    The TSF g-function of policy i applied to states x.  g = (W, b): Linear(S, G) (tsfdqn.py:537-539).  g = [flow_0, ..., (W, b)]
    with flow_k = (weight [1,S], bias [1], scale [1,S]): the normalising-flow form of tsfdqn_nf.py:331-358 -- planar flows
    z <- z + scale * tanh(z . weight + bias) followed by the Linear(S, G).
    """
    if isinstance(g, tuple):
        return torch.nn.functional.linear(x, *g)
    z = x
    for weight, bias, scale in g[:-1]:
        z = z + scale * torch.tanh(torch.nn.functional.linear(z, weight, bias))
    return torch.nn.functional.linear(z, *g[-1])


def g_flat(g):
    """Parameters of a g-function in `module.parameters()` order (flow: weight, bias, scale; then the Linear's W, b)."""
    if isinstance(g, tuple):
        return list(g)
    return [t for part in g for t in part]


def g_unflat(g, flat):
    if isinstance(g, tuple):
        return tuple(flat)
    out, k = [], 0
    for part in g:
        out.append(tuple(flat[k:k + len(part)]))
        k += len(part)
    return out


class OracleSF:
    """
    State of the SF library: N online/target psi nets, N reward maps w, per-policy Adam state (sfdqn.py:94-371),
    plus the TSF g_i / shared h (tsfdqn.py:537-560, 711-739).  Pure tensors.
    """

    def __init__(self, S, A, D, hidden=(256, 256), activations=('relu', 'relu'), lr=None, wd=None, tsf_dim=None, beta=1,
                 target_update_ev=1000, emulate=None):
        self.S, self.A, self.D = S, A, D
        self.emulate = emulate          # None: the reference's fp32 arithmetic; 'bf16' / 'tf32': tensor-core rounding points
        self.hidden, self.activations = tuple(hidden), tuple(activations)
        self.acts = make_acts(len(hidden), activations)
        self.dims = [S, hidden[0]] + list(hidden) + [A * D]      # layer l maps dims[l] -> dims[l+1]
        self.lr = dict(sf=1e-3, w=1e-3, g=1e-3, h=1e-3) if lr is None else dict(lr)
        self.wd = dict(sf=0.0, w=0.0, g=0.0, h=0.0) if wd is None else dict(wd)
        self.tsf_dim, self.beta = tsf_dim, beta
        self.target_update_ev = target_update_ev
        self.psi, self.tgt, self.w, self.g = [], [], [], []
        self.h = None
        self.adam = []                  # per policy: {'step': int, 'm': {key: tensor-list}, 'v': {...}}
        self.updates_since_target_updated = []

    @property
    def n_tasks(self):
        return len(self.psi)

    # sfdqn.py:180-213 add_training_task + :242-288 build_successor (tsfdqn.py:137-170, 227-281, 711-739)
    def add_policy(self, layers, w, g=None, h=None):
        layers = [(W.clone().float(), b.clone().float()) for W, b in layers]
        self.psi.append(layers)
        self.tgt.append([(W.clone(), b.clone()) for W, b in layers])       # update_models_weights(model, target_model)
        self.w.append(w.clone().float().reshape(1, self.D))
        if self.tsf_dim is not None:
            self.g.append(g_unflat(g, [t.clone().float() for t in g_flat(g)]))
            if self.h is None:
                self.h = (h[0].clone().float(), h[1].clone().float())
        self.adam.append({'step': 0, 'm': None, 'v': None})
        self.updates_since_target_updated.append(0)

    def to(self, device):
        """Moves every tensor of the library to `device` (bench.py's eager-GPU comparator runs this same op sequence on cuda:0)."""
        mv = lambda t: t.to(device)
        self.psi = [[(mv(W), mv(b)) for W, b in layers] for layers in self.psi]
        self.tgt = [[(mv(W), mv(b)) for W, b in layers] for layers in self.tgt]
        self.w = [mv(w) for w in self.w]
        self.g = [g_unflat(g, [mv(t) for t in g_flat(g)]) for g in self.g]
        if self.h is not None:
            self.h = tuple(mv(t) for t in self.h)
        for st in self.adam:
            for k in ('m', 'v'):
                if st[k] is not None:
                    st[k] = {key: [mv(t) for t in ts] for key, ts in st[k].items()}
        return self

    def add_random_policy(self, gen):
        layers = [init_linear(self.dims[l + 1], self.dims[l], gen) for l in range(len(self.dims) - 1)]
        w = (torch.rand(1, self.D, generator=gen) * 0.02 - 0.01)           # U(-0.01, 0.01), sfdqn.py:197
        g = h = None
        if self.tsf_dim is not None:
            g = init_linear(self.tsf_dim, self.S, gen)                    # tsfdqn.py:537-539
            if self.h is None:
                h = init_linear(self.D, self.tsf_dim, gen)                # tsfdqn.py:548-560
        self.add_policy(layers, w, g, h)

    def _mlp(self, layers, x):
        if self.emulate is None:
            return mlp_forward(layers, self.acts, x)
        return mlp_forward_emul(layers, self.acts, x, self.emulate)

    # ---------------- A2 / A3: forwards ----------------
    def get_successor(self, x, i):                                        # sfdqn.py:290-293
        return self._mlp(self.psi[i], x).reshape(-1, self.A, self.D)

    def get_successors(self, x):                                          # sfdqn.py:295-301
        return torch.stack([self.get_successor(x, i) for i in range(self.n_tasks)], dim=1)

    def get_next_successor(self, x, i):                                   # tsfdqn.py:296-299 (target nets)
        return self._mlp(self.tgt[i], x).reshape(-1, self.A, self.D)

    def get_next_successors(self, x):                                     # tsfdqn.py:301-307
        return torch.stack([self.get_next_successor(x, i) for i in range(self.n_tasks)], dim=1)

    # ---------------- A4 / A5: GPI ----------------
    def GPI_w(self, x, w):                                                # sfdqn.py:215-240
        if self.emulate is not None:
            return self._gpi_w_folded(x, w)
        psi = self.get_successors(x)                                      # [B,N,A,D]
        q = torch.nn.functional.linear(psi, w.reshape(1, self.D))[:, :, :, 0]
        task = torch.squeeze(torch.argmax(torch.max(q, dim=2).values, dim=1))
        return q, task

    def _gpi_w_folded(self, x, w, policies=None):
        """
        Emulation of the kernels' GPI form: w is folded into the output layer in fp32, Wq[a,:] = sum_d w[d] W_out[a*D+d,:]
        (bq likewise), Wq is rounded to the operand type and q = round(h_last) . Wq + bq -- psi is never formed.
        """
        qs = []
        L = len(self.acts)
        for layers in (self.psi if policies is None else [self.psi[j] for j in policies]):
            h = mlp_forward_emul(layers, self.acts, x, self.emulate, upto=L - 1)
            Wo, bo = layers[-1]
            Wq = torch.einsum('d,adk->ak', w.reshape(-1), Wo.reshape(self.A, self.D, -1))
            bq = bo.reshape(self.A, self.D) @ w.reshape(-1)
            qs.append(torch.addmm(bq, h, ROUNDERS[self.emulate](Wq).t()))
        q = torch.stack(qs, dim=1)
        task = torch.squeeze(torch.argmax(torch.max(q, dim=2).values, dim=1))
        return q, task

    def GPI(self, x, task_index):                                         # sfdqn.py:153-178
        return self.GPI_w(x, self.w[task_index])

    def next_actions(self, next_states, i, use_gpi):                      # sfdqn.py:314-322 / tsfdqn.py:604-612
        if use_gpi:
            q1, _ = self.GPI(next_states, i)
            return torch.argmax(torch.max(q1, dim=1).values, dim=-1)
        if self.emulate is not None:                                      # the kernels' folded form, own policy only
            return torch.argmax(self._gpi_w_folded(next_states, self.w[i], [i])[0][:, 0, :], dim=-1)
        sf = self.get_successor(next_states, i)
        q1 = torch.nn.functional.linear(sf, self.w[i])                    # [B,A,1]
        return torch.squeeze(torch.argmax(q1, dim=1), dim=1)

    # ---------------- A7: Adam (torch/optim/adam.py::_single_tensor_adam) ----------------
    @staticmethod
    def _adam_tensor(p, g, m, v, step, lr, wd):
        if wd != 0:
            g = g.add(p, alpha=wd)
        m.lerp_(g, 1 - ADAM_BETA1)
        v.mul_(ADAM_BETA2).addcmul_(g, g, value=1 - ADAM_BETA2)
        bc1 = 1 - ADAM_BETA1 ** step
        bc2 = 1 - ADAM_BETA2 ** step
        step_size = lr / bc1
        denom = (v.sqrt() / math.sqrt(bc2)).add_(ADAM_EPS)
        p.addcdiv_(m, denom, value=-step_size)

    def _adam_step(self, i, groups):
        """groups: list of (key, [params], [grads], lr, wd). Moments are per policy-optimizer, also for the shared h."""
        st = self.adam[i]
        if st['m'] is None:
            st['m'], st['v'] = {}, {}
        st['step'] += 1
        for key, params, grads, lr, wd in groups:
            if key not in st['m']:
                st['m'][key] = [torch.zeros_like(p) for p in params]
                st['v'][key] = [torch.zeros_like(p) for p in params]
            for p, g, m, v in zip(params, grads, st['m'][key], st['v'][key]):
                self._adam_tensor(p, g, m, v, st['step'], lr, wd)

    def _target_sync(self, i):                                            # sfdqn.py:365-369, utils/torch.py:31-33
        self.updates_since_target_updated[i] += 1
        if self.updates_since_target_updated[i] >= self.target_update_ev:
            for (Wt, bt), (W, b) in zip(self.tgt[i], self.psi[i]):
                Wt.copy_(W)
                bt.copy_(b)
            self.updates_since_target_updated[i] = 0

    # ---------------- A6: G2 train step, sfdqn.py:303-371 ----------------
    def update_successor(self, transitions, i, use_gpi=True, with_reward_loss=True):
        """with_reward_loss=False gives the G1 step of features/deep.py:93-131 (5-tuple transitions, l1 only)."""
        if transitions is None:
            return None
        if with_reward_loss:
            states, actions, rs, phis, next_states, gammas = transitions
        else:
            states, actions, phis, next_states, gammas = transitions
            rs = None
        B = len(gammas)
        idx = torch.arange(B)
        gammas = gammas.reshape(-1, 1)
        with torch.no_grad():
            next_actions = self.next_actions(next_states, i, use_gpi)
            targets = phis + gammas * self.get_next_successor(next_states, i)[idx, next_actions, :]
        flat = [t for Wb in self.psi[i] for t in Wb]
        leaves = [t.detach().requires_grad_(True) for t in flat]
        layers = [(leaves[2 * l], leaves[2 * l + 1]) for l in range(len(self.psi[i]))]
        w_leaf = self.w[i].detach().requires_grad_(True)
        cur = self._mlp(layers, states).reshape(B, self.A, self.D)
        merge = cur.clone()
        merge[idx, actions, :] = targets
        l1 = torch.nn.functional.mse_loss(cur, merge)
        if with_reward_loss:
            l2 = torch.nn.functional.mse_loss(torch.nn.functional.linear(phis, w_leaf), rs)
            loss = l1 + l2
            grads = torch.autograd.grad(loss, leaves + [w_leaf])
            groups = [('sf', flat, list(grads[:-1]), self.lr['sf'], self.wd['sf']),
                      ('w', [self.w[i]], [grads[-1]], self.lr['w'], self.wd['w'])]
        else:
            l2 = torch.zeros(())
            loss = l1
            grads = torch.autograd.grad(loss, leaves)
            groups = [('sf', flat, list(grads), self.lr['sf'], self.wd['sf'])]
        self.last_grads = {k: [g.clone() for g in gr] for k, _, gr, _, _ in groups}
        with torch.no_grad():
            self._adam_step(i, groups)
            self._target_sync(i)
        return loss.detach(), l1.detach(), l2.detach()

    # ---------------- A6': G3 TSF train step, tsfdqn.py:588-709 ----------------
    def tsf_update_successor(self, transitions, i, use_gpi=True):
        if transitions is None:
            return None
        if self.h is None:
            raise Exception('Affine Function (h) is not initialized')       # tsfdqn.py:592-593
        states, actions, rs, phis, next_states, gammas = transitions
        B = len(gammas)
        idx = torch.arange(B)
        gammas = gammas.reshape(-1, 1)
        with torch.no_grad():
            next_actions = self.next_actions(next_states, i, use_gpi)
            next_psis = self.get_next_successor(next_states, i)[idx, next_actions, :]
        flat = [t for Wb in self.psi[i] for t in Wb]
        leaves = [t.detach().requires_grad_(True) for t in flat]
        layers = [(leaves[2 * l], leaves[2 * l + 1]) for l in range(len(self.psi[i]))]
        w_leaf = self.w[i].detach().requires_grad_(True)
        g_leaf = [t.detach().requires_grad_(True) for t in g_flat(self.g[i])]
        h_leaf = [t.detach().requires_grad_(True) for t in self.h]
        lin = torch.nn.functional.linear
        cur = self._mlp(layers, states).reshape(B, self.A, self.D)
        g_fn = g_unflat(self.g[i], g_leaf)                                  # Linear (tsfdqn.py) or flow chain (tsfdqn_nf.py:653-654)
        ts, ts1 = g_apply(states, g_fn), g_apply(next_states, g_fn)         # :621-622
        aff = lin(ts, *h_leaf) + lin(ts1, *h_leaf)                          # :623
        tphis = aff * phis                                                  # :624
        targets = tphis + gammas * next_psis                                # :629 (carries grad to g_i, h)
        merge = cur.clone()
        merge[idx, actions, :] = targets                                    # :632-633 (clone is NOT detached)
        l1 = torch.nn.functional.mse_loss(cur, merge)
        l2 = torch.nn.functional.mse_loss(lin(tphis, w_leaf), rs)
        loss = l1 + torch.tensor(self.beta) * l2                            # :639-644
        grads = torch.autograd.grad(loss, leaves + [w_leaf] + g_leaf + h_leaf)
        n = len(leaves)
        groups = [('sf', flat, list(grads[:n]), self.lr['sf'], self.wd['sf']),
                  ('w', [self.w[i]], [grads[n]], self.lr['w'], self.wd['w']),
                  ('g', g_flat(self.g[i]), list(grads[n + 1:n + 1 + len(g_leaf)]), self.lr['g'], self.wd['g']),
                  ('h', list(self.h), list(grads[n + 1 + len(g_leaf):]), self.lr['h'], self.wd['h'])]
        self.last_grads = {k: [g.clone() for g in gr] for k, _, gr, _, _ in groups}
        with torch.no_grad():
            self._adam_step(i, groups)
            self._target_sync(i)
        return loss.detach(), l1.detach(), l2.detach()

    # ---------------- N1: TSF target-task adaptation, tsfdqn.py:859-997 ----------------
    def new_target_task(self, w_target, omegas, lr_omega, wd_omega=0.0, lr_omega_decay=0.0, l1_coef=0.0, gamma=0.9):
        """State of one target task as TSFDQN.train builds it (tsfdqn.py:797-832): w [1,D], omegas [1,N,1,1], one Adam."""
        return dict(w=w_target.clone(), omegas=omegas.clone(), lr_omega=lr_omega, wd_omega=wd_omega, decay=lr_omega_decay,
                    l1=l1_coef, gamma=gamma, epoch=0, step=0,
                    m=[torch.zeros_like(w_target), torch.zeros_like(omegas)], v=[torch.zeros_like(w_target), torch.zeros_like(omegas)])

    def target_q(self, s, tt):
        """Greedy branch of get_test_action (tsfdqn.py:864-871): q = w(sum_j omega_j psi_j(s)) -> [1, A, 1]."""
        norm = tt['omegas'] / torch.sum(tt['omegas'], dim=1, keepdim=True)
        tsf = torch.sum(self.get_successors(s) * norm, dim=1)
        return torch.nn.functional.linear(tsf, tt['w'])

    def target_adapt_step(self, tt, s, a, r, s1, a1, phi):
        """update_test_reward_mapper (tsfdqn.py:917-997) + scheduler.step() (:895): returns (loss, l2, l1)."""
        lin = torch.nn.functional.linear
        w_leaf, om_leaf = tt['w'].detach().requires_grad_(True), tt['omegas'].detach().requires_grad_(True)
        norm = om_leaf / torch.sum(om_leaf, dim=1, keepdim=True)                               # :929
        with torch.no_grad():
            ts = torch.vstack([g_apply(s, g) for g in self.g]).unsqueeze(1)                   # :931-940
            ts1 = torch.vstack([g_apply(s1, g) for g in self.g]).unsqueeze(1)
            psi, next_psi = self.get_successors(s), self.get_next_successors(s1)               # :948-950
        aff = lin(torch.sum(ts * norm, dim=1), *self.h) + lin(torch.sum(ts1 * norm, dim=1), *self.h)     # :942-944
        tphi = phi * aff.squeeze(0)                                                          # :945
        next_tsf = tphi + tt['gamma'] * torch.sum(next_psi * norm, dim=1)[:, a1, :]          # :953
        tsf = torch.sum(psi * norm, dim=1)[:, a, :]                                          # :955
        l1 = torch.nn.functional.mse_loss(tsf, next_tsf)
        l2 = torch.mean((lin(tphi, w_leaf) - torch.tensor([r], dtype=torch.float32)) ** 2)   # [1,1] vs [1] broadcast, :958-969
        loss = l1 + torch.tensor(self.beta) * l2 + torch.tensor(tt['l1']) * torch.norm(om_leaf, 1)   # :962-971
        gw, go = torch.autograd.grad(loss, [w_leaf, om_leaf])
        with torch.no_grad():
            tt['step'] += 1
            lr_o = tt['lr_omega'] * (1 - tt['decay']) ** tt['epoch']                         # LambdaLR on the omega group (:822-826)
            self._adam_tensor(tt['w'], gw, tt['m'][0], tt['v'][0], tt['step'], self.lr['w'], self.wd['w'])
            self._adam_tensor(tt['omegas'], go, tt['m'][1], tt['v'][1], tt['step'], lr_o, tt['wd_omega'])
            tt['omegas'].clamp_(1e-7)                                                        # :977-979
            tt['epoch'] += 1
        return loss.detach(), l2.detach(), l1.detach()

    # ---------------- A6'': G1 ensemble step, agents/sfdqn.py:57-60 ----------------
    def ensemble_update_sequential(self, transitions5):
        """Literal reference loop (Gauss-Seidel: task k's GPI sees psi_0..psi_{k-1} already stepped)."""
        return [self.update_successor(transitions5, i, True, with_reward_loss=False) for i in range(self.n_tasks)]

    def ensemble_update_frozen(self, transitions, tsf=False, use_gpi=True, with_reward_loss=True):
        """
        Frozen-snapshot (Jacobi) ensemble step, SURVEY.md section 8c: every policy i is updated by the reference step
        computed from the SAME pre-step library.  Implemented with reference-shaped steps only: deepcopy the library
        once per task, step copy i, and collect policy i from copy i.
        For TSF the shared h is stepped once per policy-optimizer, applied in policy order (h <- h + sum_i delta_i,
        every delta_i computed from the pre-step h and optimizer i's own moments).
        """
        snaps, outs = [], []
        for i in range(self.n_tasks):
            c = copy.deepcopy(self)
            if tsf:
                outs.append(c.tsf_update_successor(transitions, i, use_gpi))
            else:
                outs.append(c.update_successor(transitions, i, use_gpi, with_reward_loss))
            snaps.append(c)
        h0 = None if self.h is None else [t.clone() for t in self.h]
        for i, c in enumerate(snaps):
            self.psi[i], self.tgt[i], self.w[i], self.adam[i] = c.psi[i], c.tgt[i], c.w[i], c.adam[i]
            self.updates_since_target_updated[i] = c.updates_since_target_updated[i]
            if tsf:
                self.g[i] = c.g[i]
                for t, t0, tc in zip(self.h, h0, c.h):
                    t.add_(tc - t0)
        return outs


# ----------------------------------------------------------------------------------------------------------------
# N3: learned features -- PhiFunction (sfdqn_phi.py:90-123) and SFDQN.pre_train (sfdqn_phi.py:800-873)
# ----------------------------------------------------------------------------------------------------------------
class OraclePhi:
    """phi_theta = Linear(2S + action_dim, 128) - ReLU - Linear(128, 256) - ReLU - Linear(256, D), one Adam(lr = 1e-3)."""

    ACTS = ('relu', 'relu', 'none')

    def __init__(self, layers, alpha_phi=1e-3):
        self.layers = [(W.clone().float(), b.clone().float()) for W, b in layers]
        self.alpha_phi = alpha_phi
        self.adam = {'step': 0, 'm': [torch.zeros_like(t) for Wb in self.layers for t in Wb],
                     'v': [torch.zeros_like(t) for Wb in self.layers for t in Wb]}

    @staticmethod
    def inputs(state, action, next_state):                               # PhiFunction.forward, sfdqn_phi.py:106-117
        if action.ndim == 0:
            action = action.unsqueeze(0)
        if action.ndim == 1:
            action = action.unsqueeze(1)
        if state.ndim == 1:
            state = state.unsqueeze(0)
        if next_state.ndim == 1:
            next_state = next_state.unsqueeze(0)
        return torch.cat([state, action, next_state], axis=1)

    def forward(self, state, action, next_state):
        return mlp_forward(self.layers, self.ACTS, self.inputs(state, action, next_state))

    def regression_step(self, state, action, reward, next_state, head):
        """Body of pre_train's inner loop (sfdqn_phi.py:850-866); head = {'w': [1,D], 'm', 'v', 'step'} (fit_w + its Adam)."""
        flat = [t for Wb in self.layers for t in Wb]
        leaves = [t.detach().requires_grad_(True) for t in flat]
        layers = [(leaves[2 * l], leaves[2 * l + 1]) for l in range(len(self.layers))]
        w_leaf = head['w'].detach().requires_grad_(True)
        phis = mlp_forward(layers, self.ACTS, self.inputs(state, action, next_state))
        lin = torch.nn.functional.linear(phis, w_leaf)
        loss = torch.nn.functional.mse_loss(reward, lin)                 # :859 (reward [B,1], lin [B,1])
        grads = torch.autograd.grad(loss, leaves + [w_leaf])
        with torch.no_grad():
            self.adam['step'] += 1
            for p, g, m, v in zip(flat, grads[:-1], self.adam['m'], self.adam['v']):
                OracleSF._adam_tensor(p, g, m, v, self.adam['step'], self.alpha_phi, 0.0)
            head['step'] += 1
            OracleSF._adam_tensor(head['w'], grads[-1], head['m'], head['v'], head['step'], 1e-3, 0.0)
        return float(loss.detach())


def oracle_phi_pre_train(phi, head_ws, train_tasks, n_samples_pre_train, n_cycles=5, n_batch=32):
    """
    SFDQN.pre_train (sfdqn_phi.py:800-873) with the initial weights handed in (phi: OraclePhi, head_ws: list of [1,D]).
    Restates the loop and the replay ring (ReplayBuffer, sfdqn_phi.py:9-86: picks = np.random.randint(0, size, n_batch) once
    per update; actions from random.randrange) with plain lists / tensors.  Returns the list of losses.
    """
    import random
    import numpy as np
    heads = [dict(w=w.clone().float(), m=torch.zeros_like(w), v=torch.zeros_like(w), step=0) for w in head_ws]
    ring, losses = [], []
    n_actions = train_tasks[0].action_count()
    for cycle in range(n_cycles):
        for task_id, task in enumerate(train_tasks):
            s_enc = task.encode(task.initialize())
            for sample in range(n_samples_pre_train):
                a = random.randrange(n_actions)
                s1, r, terminal = task.transition(a)
                s1_enc = task.encode(s1)
                ring.append((s_enc, torch.tensor(a), torch.tensor(r, dtype=torch.float32).float(), s1_enc))
                s_enc = s1_enc
                if terminal:
                    s_enc = task.encode(task.initialize())
                if len(ring) >= n_batch:
                    picks = np.random.randint(low=0, high=len(ring), size=(n_batch,))
                    st, ac, rw, ns = zip(*[ring[k] for k in picks])
                    losses.append(phi.regression_step(torch.vstack(st), torch.tensor(ac), torch.vstack(rw), torch.vstack(ns),
                                                      heads[task_id]))
    phi.heads = heads
    return losses


# ----------------------------------------------------------------------------------------------------------------
# G4: joint psi / phi step -- features/deep_phi.py:95-224 (DeepSF_PHI.update_successor) with the shared phi model and the
# per-task loss coefficient of agents/sfdqn_phi.py:144-165.  ORACLE ONLY so far (second half): the product does
# not implement this step yet; the restatement is pinned to the reference so that the kernel path can be built against it.
# ----------------------------------------------------------------------------------------------------------------
class OracleG4:
    """
    State: N psi nets + targets (as OracleSF), reward maps fit_w[i] = Linear(D, 1) WITH bias (features/deep_phi.py:268),
    one shared phi MLP over cat[s, a, s'] (main_sfdqn_phi_torch.py:52-73), one loss coefficient per task (init 1).
    Every call builds a FRESH torch.optim.Adam (features/deep_phi.py:170), so each update is a first Adam step
    (step = 1, zero moments: p -= lr * g / (|g| + eps) up to rounding); the coefficient's group has maximize=True.
    """

    LR = 1e-3

    def __init__(self, A, D, psi_acts, phi_layers, phi_acts, target_update_ev=1000):
        self.A, self.D = A, D
        self.psi_acts, self.phi_acts = list(psi_acts), list(phi_acts)
        self.phi = [(W.clone().float(), b.clone().float()) for W, b in phi_layers]
        self.psi, self.tgt, self.w, self.coef = [], [], [], []
        self.target_update_ev = target_update_ev
        self.updates_since_target_updated = []

    def add_policy(self, layers, w):
        self.psi.append([(W.clone().float(), b.clone().float()) for W, b in layers])
        self.tgt.append([(W.clone(), b.clone()) for W, b in self.psi[-1]])
        self.w.append((w[0].clone().float(), w[1].clone().float()))          # (weight [1,D], bias [1])
        self.coef.append(torch.ones(1))
        self.updates_since_target_updated.append(0)

    def get_successor(self, x, i, layers=None):
        return mlp_forward(self.psi[i] if layers is None else layers, self.psi_acts, x).reshape(-1, self.A, self.D)

    def GPI(self, x, i):                                                      # features/deep_phi.py:226-251: q = w(psi)[..., 0]
        psi = torch.stack([self.get_successor(x, j) for j in range(len(self.psi))], dim=1)
        q = torch.nn.functional.linear(psi, *self.w[i])[:, :, :, 0]
        return q, torch.squeeze(torch.argmax(torch.max(q, dim=2).values, dim=1))

    def update_successor(self, transitions, i, use_gpi=True):
        if transitions is None:
            return None
        states, actions, rs, _, next_states, gammas = transitions
        B = len(gammas)
        idx = torch.arange(B)
        gammas = gammas.reshape(-1, 1)
        lin = torch.nn.functional.linear
        flat_phi = [t for Wb in self.phi for t in Wb]
        flat_psi = [t for Wb in self.psi[i] for t in Wb]
        leaves_phi = [t.detach().requires_grad_(True) for t in flat_phi]
        leaves_psi = [t.detach().requires_grad_(True) for t in flat_psi]
        w_leaf = [t.detach().requires_grad_(True) for t in self.w[i]]
        c_leaf = self.coef[i].detach().requires_grad_(True)
        phi_layers = [(leaves_phi[2 * l], leaves_phi[2 * l + 1]) for l in range(len(self.phi))]
        psi_layers = [(leaves_psi[2 * l], leaves_psi[2 * l + 1]) for l in range(len(self.psi[i]))]
        x_phi = torch.cat([states, actions.reshape(B, 1), next_states], dim=1)               # :110
        phis = mlp_forward(phi_layers, self.phi_acts, x_phi)                                  # :111 (carries grad)
        with torch.no_grad():                                                                 # next actions: indices only
            if use_gpi:
                q1, _ = self.GPI(next_states, i)                                              # :114-116
                next_actions = torch.argmax(torch.max(q1, dim=1).values, dim=-1)
            else:
                q1 = lin(self.get_successor(next_states, i), *self.w[i])                      # :118-123
                next_actions = torch.squeeze(torch.argmax(q1, dim=1), dim=1)
            next_psi = mlp_forward(self.tgt[i], self.psi_acts, next_states).reshape(B, self.A, self.D)[idx, next_actions, :]
        cur = self.get_successor(states, i, psi_layers)
        targets = phis + gammas * next_psi                                                    # :136 (phi's grad flows in)
        merge = cur.clone()
        merge[idx, actions, :] = targets                                                      # :142-143
        r_fit = lin(phis, *w_leaf)                                                            # :147
        phi_loss = torch.nn.functional.mse_loss(r_fit, rs).unsqueeze(0)                       # :175
        psi_loss = torch.nn.functional.mse_loss(cur, merge).unsqueeze(0)                      # :180
        loss = phi_loss + c_leaf * psi_loss                                                   # :185
        params = leaves_psi + leaves_phi + w_leaf + [c_leaf]
        grads = torch.autograd.grad(loss, params)
        with torch.no_grad():
            targets_flat = flat_psi + flat_phi + list(self.w[i]) + [self.coef[i]]
            for k, (p, g) in enumerate(zip(targets_flat, grads)):
                g = g.clamp(-1e10, 1e10)                                                      # :205-208
                if k == len(targets_flat) - 1:
                    g = -g                                                                    # maximize=True (:169)
                OracleSF._adam_tensor(p, g, torch.zeros_like(p), torch.zeros_like(p), 1, self.LR, 0.0)
            self.coef[i].clamp_(1e-2, 1e6)                                                    # :212-215
            self.updates_since_target_updated[i] += 1
            if self.updates_since_target_updated[i] >= self.target_update_ev:                 # :219-224
                for (Wt, bt), (W, b) in zip(self.tgt[i], self.psi[i]):
                    Wt.copy_(W)
                    bt.copy_(b)
                self.updates_since_target_updated[i] = 0
        return loss.detach(), psi_loss.detach(), phi_loss.detach(), self.coef[i].clone()


# ----------------------------------------------------------------------------------------------------------------
# Synthetic replay batches (SURVEY.md section 8d "Synthetic inputs")
# ----------------------------------------------------------------------------------------------------------------
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
