"""torchrun script: per-phase CUDA-event timing of the sharded train step (which collective / segment costs what)."""
import os, sys, time
os.environ["SFGPI_PEER"] = "0"      # this script times the NCCL-collective variant of the sharded step phase by phase
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from tests.synthetic import synthetic_transitions
import bench
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.library import _stream
from deep_successor_features_for_transfer_b200.dist import allreduce_max_keys

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfg = bench.WORKLOADS['tsfdqn_reacher_b4096']
dsf, ag = bench.build_agent(cfg, cfg['n_local'], 'bf16')
lib = dsf._library
lib.enable_sharding()
gen = torch.Generator().manual_seed(1)
trs = [tuple(t.cuda() for t in synthetic_transitions(cfg['B'], cfg['S'], cfg['A'], cfg['D'], gen)) for _ in range(4)]
for k in range(5):
    ag.update_successor_all(trs[k % 4], use_gpi=True)
plan = lib._ws[lib.last_plan_key]
segs, keys = plan['segments'], plan['keys']
st = _stream()
names = ['seg0(pack,fill)', 'gather_w', 'seg1(fold,fwd)', 'allreduce_keys', 'h0.clone', 'seg2(td,bwd,adam)', 'h delta allreduce']
acc = [0.0] * len(names)
wall = 0.0
R = 30
for k in range(R):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    ev[0].record()
    _lib.run(segs[0][0], segs[0][1], st, segs[0][2]); ev[1].record()
    lib._gather_w(plan['w_all']); ev[2].record()
    _lib.run(segs[1][0], segs[1][1], st, segs[1][2]); ev[3].record()
    allreduce_max_keys(keys, lib.shard.group); ev[4].record()
    h0 = lib.h.clone(); ev[5].record()
    _lib.run(segs[2][0], segs[2][1], st, segs[2][2]); ev[6].record()
    delta = lib.h - h0
    dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=lib.shard.group)
    lib.h.copy_(h0 + delta); ev[7].record()
    torch.cuda.synchronize()
    wall += time.perf_counter() - t0
    for i in range(len(names)):
        acc[i] += ev[i].elapsed_time(ev[i + 1])
if rank == 0:
    print(f'world={world}: ' + ' | '.join(f'{n} {a / R * 1e3:.0f}us' for n, a in zip(names, acc)) + f' | total {sum(acc) / R * 1e3:.0f}us, wall {wall / R * 1e6:.0f}us')
dist.destroy_process_group()
