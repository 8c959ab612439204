"""Kernel windows (SFGPI_TRACE) of the end-to-end step: batch pulled from pinned host memory by the prologue kernel, losses copied back."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.workloads import synthetic_transitions
cfg = bench.WORKLOADS['tsfdqn_reacher_b4096']
dsf, ag = bench.build_agent(cfg, 4, 'bf16')
gen = torch.Generator().manual_seed(1)
pinned = [tuple(t.pin_memory() for t in synthetic_transitions(4096, 4, 9, 12, gen)) for _ in range(4)]
losses_host = torch.zeros(4, 3).pin_memory()
st = torch.cuda.current_stream()
for k in range(20):
    ag.update_successor_all(pinned[k % 4], use_gpi=True, host_losses=losses_host); st.synchronize()
L = _lib.lib()
L.sfgpi_trace_enable(1)
for k in range(4):
    t0 = time.perf_counter()
    ag.update_successor_all(pinned[k % 4], use_gpi=True, host_losses=losses_host)
    t1 = time.perf_counter()
    st.synchronize()
    t2 = time.perf_counter()
    print(f'host call {1e6 * (t1 - t0):.1f} us, until synchronised {1e6 * (t2 - t0):.1f} us', file=sys.stderr)
    L.sfgpi_trace_dump()
