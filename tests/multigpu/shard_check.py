"""
Multi-GPU check (torchrun, NCCL): a policy-sharded library must reproduce the single-GPU library.
Every rank builds (a) the FULL N-policy TSF agent and (b) its own shard (policies [lo, hi)), runs K all-task train steps with
GPI on identical batches, and compares losses / weights / GPI keys of its shard against the same policies of the full run.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu/shard_check.py [fp32|bf16]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from oracle.sf_oracle import OracleSF, synthetic_transitions
from tests import gpu_util as gu
from deep_successor_features_for_transfer_b200.dist import shard_range

precision = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
S, A, D, B, K = 4, 9, 12, 1000, 3
# uneven shards (plain gather / all-reduce path) by default; 'even' exercises the fused one-all-gather exchange
N = 2 * world + (0 if (len(sys.argv) > 2 and sys.argv[2] == 'even') else 1)
gen = torch.Generator().manual_seed(5)
o = OracleSF(S, A, D, (256, 256), ('relu', 'relu'), tsf_dim=100, beta=1)
for _ in range(N):
    o.add_random_policy(gen)
batches = [synthetic_transitions(B, S, A, D, gen) for _ in range(K)]
meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N, gdim=100, beta=1, use_gpi=True)

full_sf, full_ag = gu.build_g3(meta, oracle=o)
full_sf._library.set_precision(precision)
lo, hi = shard_range(N, world, rank)


class Sub:                                                    # oracle view holding only this rank's policies
    pass


sub = Sub()
sub.psi, sub.w, sub.g, sub.h = o.psi[lo:hi], o.w[lo:hi], o.g[lo:hi], o.h
sh_sf, sh_ag = gu.build_g3(dict(meta, N=hi - lo), oracle=sub)
sh_sf._library.set_precision(precision)
sh_sf._library.enable_sharding()

# identical weights here, so the all-reduced packed keys must be BIT-equal to the single-GPU keys
x = batches[0][4].cuda()
_, ka_f, kt_f = full_sf._library.gpi(x, full_sf.fit_w[0].weight, want_q=False)
w0 = full_sf.fit_w[0].weight.detach().clone()
_, ka_s, kt_s = sh_sf._library.gpi(x, w0, want_q=False)
keys_equal = bool(torch.equal(ka_f, ka_s) and torch.equal(kt_f, kt_s))

worst = 0.0
for k in range(K):
    tr = tuple(t.cuda() for t in batches[k])
    lf = full_ag.update_successor_all(tr, use_gpi=True).cpu()
    ls = sh_ag.update_successor_all(tr, use_gpi=True).cpu()
    err = float((lf[lo:hi] - ls).abs().max() / lf.abs().max())
    worst = max(worst, err)
    for i in range(lo, hi):
        for (Wf, bf), (Ws, bs) in zip(gu.psi_params(full_sf, i), gu.psi_params(sh_sf, i - lo)):
            worst = max(worst, float((Wf - Ws).abs().max() / Wf.abs().max()))
    hf, hs = full_ag.h_function.weight.data.cpu(), sh_ag.h_function.weight.data.cpu()
    worst = max(worst, float((hf - hs).abs().max() / hf.abs().max()))
# training: the shared h receives the ranks' deltas in a different summation order -> fp32 rounding-level drift only
tol = 5e-5 if precision == 'fp32' else 5e-3
t = torch.tensor([worst, 0.0 if keys_equal else 1.0], device='cuda')
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ok = float(t[0]) < tol and float(t[1]) == 0.0
    print(f'shard_check {precision} world={world} N={N}: max rel deviation {float(t[0]):.3e} (tol {tol}), '
          f'GPI keys equal: {float(t[1]) == 0.0} -> {"OK" if ok else "FAIL"}', flush=True)
dist.destroy_process_group()
