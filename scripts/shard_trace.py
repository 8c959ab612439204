"""torchrun script: kernel-window trace (SFGPI_TRACE=1) of the policy-sharded train step on rank 0, peer-memory transport."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.workloads import synthetic_transitions

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg = bench.WORKLOADS['tsfdqn_reacher_b4096']
dsf, ag = bench.build_agent(cfg, cfg['n_local'], 'bf16', first_policy=rank * cfg['n_local'])
lib = dsf._library
lib.enable_sharding()
gen = torch.Generator().manual_seed(1)
trs = [tuple(t.cuda() for t in synthetic_transitions(cfg['B'], 4, 9, 12, gen)) for _ in range(4)]
for k in range(50):
    ag.update_successor_all(trs[k % 4], use_gpi=True)
torch.cuda.synchronize(); dist.barrier()
L = _lib.lib()
L.sfgpi_trace_enable(1)
for k in range(6):
    ag.update_successor_all(trs[k % 4], use_gpi=True)
    if rank == 0 and k >= 3:
        L.sfgpi_trace_dump()
    else:
        torch.cuda.synchronize()
    dist.barrier()
dist.destroy_process_group()
