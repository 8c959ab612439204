"""
Agent-level parity helpers (SURVEY section 8b "what calls it", VERDICT r1 "agent-level drop-in"): the loop body of the
reference's train() -- per task set_active_training_task + n_samples x next_sample (sfdqn.py:550-627, 629-677;
tsfdqn.py:435-497) -- restated over the CPU oracle, so the committed agent fixtures (tests/golden/agent_*.npz, produced by the
unmodified reference agents) pin the oracle on CPU, and the same fixtures then check the CUDA agents on the GPU.
"""
import random

import numpy as np
import torch

from tests.golden_util import n_layers, t
from tests.toy_task import ToyTask


def oracle_agent_run(meta, z):
    """Returns dict(actions, losses, counters, oracle) of the oracle driven through the reference agent's loop."""
    from oracle.sf_oracle import OracleSF
    tsf = meta['kind'] == 'agent_tsfdqn'
    S, A, D, N = meta['S'], meta['A'], meta['D'], meta['N']
    o = OracleSF(S, A, D, meta['hidden'], meta['acts'], tsf_dim=meta['gdim'] if tsf else None, beta=meta['beta'],
                 target_update_ev=meta['target_update_ev'])
    for i in range(N):
        layers = [(t(z[f'init.psi{i}.W{l}']), t(z[f'init.psi{i}.b{l}'])) for l in range(n_layers(meta))]
        g = (t(z[f'init.g{i}.W']), t(z[f'init.g{i}.b'])) if tsf else None
        h = (t(z['init.h.W']), t(z['init.h.b'])) if tsf else None
        o.add_policy(layers, t(z[f'init.w{i}']), g, h)
    tasks = [ToyTask(S, A, D, i, seed=meta['seed']) for i in range(N)]
    counters = [np.zeros(N, dtype=np.int64) for _ in range(N)]
    random.seed(meta['seed'])
    np.random.seed(meta['seed'])
    losses, n_batch, T = [], meta['n_batch'], meta['T']
    for index, task in enumerate(tasks):
        ring, new_episode, steps_ep, eps, s = [], True, 0, meta['epsilon'], None
        for _ in range(meta['n_samples']):
            if new_episode:                                                    # sfdqn.py:565-575
                s, new_episode, steps_ep = task.initialize(), False, 0

            def greedy():                                                      # sfdqn.py:585-594
                q, c = o.GPI(s, index)
                if meta['use_gpi']:
                    counters[index][int(c)] += 1
                else:
                    c = index
                return int(torch.argmax(q[:, int(c), :].flatten()))
            if tsf:                                                            # tsfdqn.py:453-457: GPI before the coin
                g_a = greedy()
                a = random.randrange(A) if random.random() <= eps else g_a
            else:
                a = random.randrange(A) if random.random() <= eps else greedy()
            s1, r, terminal = task.transition(torch.tensor(a))
            gamma = 0.0 if terminal else meta['gamma']
            new_episode = new_episode or terminal
            ring.append((s, a, r.float(), task.features(s, a, s1), s1, gamma))      # ReplayBuffer.append, sfdqn.py:86-89
            if len(ring) >= n_batch:                                           # ReplayBuffer.replay, sfdqn.py:57-66
                picks = np.random.randint(low=0, high=len(ring), size=(n_batch,))
                st, ac, rw, ph, ns, gm = zip(*[ring[k] for k in picks])
                tr = (torch.vstack(st), torch.tensor(ac), torch.vstack(rw), torch.vstack(ph), torch.vstack(ns), torch.tensor(gm))
                res = o.tsf_update_successor(tr, index, meta['use_gpi']) if tsf else o.update_successor(tr, index, meta['use_gpi'])
                losses.append([float(x) for x in res])
            else:
                losses.append([float('nan')] * 3)
            s = s1
            steps_ep += 1
            if steps_ep >= T:
                new_episode = True
    actions = np.array([a for task in tasks for a in task.actions_taken], dtype=np.int64)
    return dict(actions=actions, losses=np.array(losses), counters=counters, oracle=o)
