"""Helpers shared by the parity tests: load a committed reference fixture into the CPU oracle."""
import ast
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'), allow_pickle=False)
    meta = ast.literal_eval(str(z['meta']))
    return meta, {k: z[k] for k in z.files if k != 'meta'}


def t(a):
    return torch.from_numpy(np.asarray(a).copy())


def n_layers(meta):
    return len(meta['hidden']) + 2


def net_layers(z, prefix, meta):
    return [(t(z[f'{prefix}.W{l}']), t(z[f'{prefix}.b{l}'])) for l in range(n_layers(meta))]


def oracle_from_golden(meta, z, **kw):
    from oracle.sf_oracle import OracleSF          # (lazy: bench.py's measured arm imports this module for FakeTask & co only)
    tsf = meta['kind'] in ('g3', 'g3_target')
    o = OracleSF(meta['S'], meta['A'], meta['D'], meta['hidden'], meta['acts'],
                 tsf_dim=meta.get('gdim') if tsf else None, beta=meta.get('beta', 1),
                 target_update_ev=meta.get('target_update_ev', 1000), **kw)
    for i in range(meta['N']):
        w = t(z[f'init.w{i}']).reshape(1, meta['D'])
        g = (t(z[f'init.g{i}.W']), t(z[f'init.g{i}.b'])) if tsf else None
        h = (t(z['init.h.W']), t(z['init.h.b'])) if tsf else None
        o.add_policy(net_layers(z, f'init.psi{i}', meta), w, g, h)
    return o


def transitions(z, k, five=False):
    names = ['states', 'actions', 'phis', 'next_states', 'gammas'] if five else \
        ['states', 'actions', 'rs', 'phis', 'next_states', 'gammas']
    return tuple(t(z[f'tr{k}.{n}']) for n in names)


def rel_err(a, b):
    """max |a-b| / max |b|  (scale-normalised max error; the parity metric used throughout the tests)."""
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
