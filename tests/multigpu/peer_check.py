"""
Multi-GPU check + timing of the peer-memory exchange (torchrun, one rank per GPU):
two policy-sharded TSF agents are built from the same weights on every rank, one exchanging through the peer-memory kernels
(csrc/peer.cu), one through the NCCL collectives (SFGPI_PEER=0).  Both exchanges are exact (integer MAX of keys; the h deltas
are summed in rank order either way), so after K train steps -- with deliberate rank skew to shake the flag protocol -- the
two agents must be BIT-identical.  Then both are timed with CUDA events.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/peer_check.py [fp32|bf16] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from oracle.sf_oracle import OracleSF, synthetic_transitions
from tests import gpu_util as gu
from deep_successor_features_for_transfer_b200.dist import shard_range

precision = sys.argv[1] if len(sys.argv) > 1 else 'bf16'
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
S, A, D, B = 4, 9, 12, 4096
N = 4 * world
gen = torch.Generator().manual_seed(11)
o = OracleSF(S, A, D, (256, 256), ('relu', 'relu'), tsf_dim=100, beta=1)
for _ in range(N):
    o.add_random_policy(gen)
batches = [tuple(t.cuda() for t in synthetic_transitions(B, S, A, D, gen)) for _ in range(4)]
meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N, gdim=100, beta=1, use_gpi=True)
lo, hi = shard_range(N, world, rank)


class Sub:
    pass


def build(peer):
    sub = Sub()
    sub.psi, sub.w, sub.g, sub.h = o.psi[lo:hi], o.w[lo:hi], o.g[lo:hi], o.h
    sf, ag = gu.build_g3(dict(meta, N=hi - lo), oracle=sub)
    sf._library.set_precision(precision)
    os.environ['SFGPI_PEER'] = '1' if peer else '0'
    sf._library.enable_sharding()
    return sf, ag


sf_p, ag_p = build(True)
sf_n, ag_n = build(False)
peer_on = sf_p._library._peer is not None
assert sf_n._library._peer is None

same = True
for k in range(K):
    if k % 7 == rank % 7:
        time.sleep(0.01 * (1 + rank))                        # rank skew: peers arrive at the exchange at different times
    lp = ag_p.update_successor_all(batches[k % 4], use_gpi=True)
    ln = ag_n.update_successor_all(batches[k % 4], use_gpi=True)
    same = same and bool(torch.equal(lp, ln))
torch.cuda.synchronize()
lib_p, lib_n = sf_p._library, sf_n._library
n = hi - lo
for a, b in ((lib_p.online[:n], lib_n.online[:n]), (lib_p.w[:n], lib_n.w[:n]), (lib_p.g[:n], lib_n.g[:n]), (lib_p.h, lib_n.h),
             (lib_p._xchg['w_all'], lib_n._xchg['w_all'])):
    same = same and bool(torch.equal(a, b))


def timed(ag, steps=50):
    for k in range(5):
        ag.update_successor_all(batches[k % 4], use_gpi=True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        ag.update_successor_all(batches[k % 4], use_gpi=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


t_p, t_n = timed(ag_p), timed(ag_n)
flag = torch.tensor([0.0 if same else 1.0, 0.0 if peer_on else 1.0], device='cuda')
dist.all_reduce(flag, op=dist.ReduceOp.MAX)
if rank == 0:
    ok = float(flag[0]) == 0.0 and float(flag[1]) == 0.0
    print(f'peer_check {precision} world={world} N={N} B={B} K={K}: peer mode active: {float(flag[1]) == 0.0}, bit-identical to the '
          f'NCCL exchange: {float(flag[0]) == 0.0} -> {"OK" if ok else "FAIL"}; step (back to back, max over ranks): '
          f'peer {t_p * 1e3:.1f} us, NCCL {t_n * 1e3:.1f} us', flush=True)
dist.destroy_process_group()
