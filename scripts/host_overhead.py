"""Host-side cost of one train-step call (python + ctypes + launches), measured while the GPU queue is short."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tests.synthetic import synthetic_transitions

cfg = bench.WORKLOADS['tsfdqn_reacher_b4096']
torch.cuda.set_device(0)
dsf, ag = bench.build_agent(cfg, cfg['n_local'], 'bf16')
gen = torch.Generator().manual_seed(1)
host = [tuple(t.pin_memory() for t in synthetic_transitions(cfg['B'], cfg['S'], cfg['A'], cfg['D'], gen)) for _ in range(4)]
dev = [tuple(t.cuda() for t in tr) for tr in host]
for k in range(20):
    ag.update_successor_all(dev[k % 4], use_gpi=True)
    ag.update_successor_all(host[k % 4], use_gpi=True)
for name, data in (('device inputs', dev), ('pinned host inputs', host)):
    tot = 0.0
    for rep in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(10):
            ag.update_successor_all(data[k % 4], use_gpi=True)
        tot += time.perf_counter() - t0
    print(f'{name}: {tot / 200 * 1e6:.1f} us of host time per call', flush=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(200):
    l = ag.update_successor_all(host[k % 4], use_gpi=True).cpu()
print(f'e2e loop: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us per step', flush=True)
pr = cProfile.Profile()
pr.enable()
for k in range(300):
    ag.update_successor_all(host[k % 4], use_gpi=True)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
