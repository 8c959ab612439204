"""cProfile of the host side of the end-to-end train step (pinned host batches, losses read back)."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deep_successor_features_for_transfer_b200.workloads import synthetic_transitions
cfg = bench.WORKLOADS['tsfdqn_reacher_b4096']
dsf, ag = bench.build_agent(cfg, 4, 'bf16')
gen = torch.Generator().manual_seed(1)
pinned = [tuple(t.pin_memory() for t in synthetic_transitions(4096, 4, 9, 12, gen)) for _ in range(4)]
losses_host = torch.zeros(4, 3).pin_memory()
st = torch.cuda.current_stream()
for k in range(50):
    ag.update_successor_all(pinned[k % 4], use_gpi=True, host_losses=losses_host); st.synchronize()
N = 300
t0 = time.perf_counter()
for k in range(N):
    ag.update_successor_all(pinned[k % 4], use_gpi=True, host_losses=losses_host)
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f'host call (no sync between): {1e6 * (t1 - t0) / N:.1f} us')
pr = cProfile.Profile()
pr.enable()
for k in range(N):
    ag.update_successor_all(pinned[k % 4], use_gpi=True, host_losses=losses_host)
pr.disable()
torch.cuda.synchronize()
ps = pstats.Stats(pr).sort_stats('tottime')
ps.print_stats(18)
