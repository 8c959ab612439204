"""
Synthetic replay batches of the reference's shapes and value ranges (SURVEY 8d "Synthetic inputs") for the benchmark, the
measurement scripts and the GPU tests.  Deliberately NOT under oracle/: the measured arm of bench.py must not import the
checker.  (oracle/sf_oracle.py carries the same generator for its own tests; tests/test_oracle_golden.py pins the two to
each other.)
"""
import torch


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
