# -*- coding: UTF-8 -*-
"""
bench.py -- headline benchmark of the SF/GPI hot path (contract: one JSON line on stdout, rank 0).

Workload (BASELINE.json configs[1]): TSFDQN on Reacher shapes (S=4, A=9, D=12, MLP 256-256
relu, g: 4->100, h: 100->12, beta=1), synthetic replay batch B=4096, 4 source policies PER GPU, GPI next actions.
One "step" = one replay batch on which EVERY policy of the library is updated (fused all-task TD update, frozen-snapshot
semantics) => transitions x tasks = B * N_total SF TD updates per step.  Multi-GPU: policies sharded 4 per GPU (weak
scaling), the batch replicated, GPI's max over policies exchanged as packed int64 keys (NCCL MAX all-reduce).

  value : updates/s with the batches already resident in HBM (CUDA events per step, L2 flushed between steps)
  e2e   : same metric through the public API (TSFDQN.update_successor_all) from pinned HOST batches, H2D copies and the
          D2H read of the losses inside the timed region (wall clock, synchronised both sides)
  --impl reference : the reference algorithm on the host cores (CPU oracle port, torch CPU fp32, all threads): the same
          all-task update done the way the reference does it, one update_successor call per task on the same batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    'tsfdqn_reacher_b4096': dict(S=4, A=9, D=12, hidden=[256, 256], acts=['relu', 'relu'], gdim=100, beta=1, B=4096,
                                 n_local=4, hopper=False),
    # BASELINE config 4 (ii): TSFDQN dissimilar-task sequence, 256 policies in total split over the GPUs (strong scaling),
    # beta = 30 (reacher_dissimilar.cfg:40), every policy stepped on every batch with GPI over all 256
    'tsfdqn_dissimilar_n256': dict(S=4, A=9, D=12, hidden=[256, 256], acts=['relu', 'relu'], gdim=100, beta=30, B=4096,
                                   n_total=256, hopper=False),
}
SEED = 1024
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the ncu --set full capture of this
# command (profiles/): cold-cache replay, so it is an upper bound of the in-step traffic
KERNEL_DRAM_BYTES = {'bf16': 4256000, 'fp32': None}      # profiles/r01_step_kernels_full_v2.md: 3.940 MB read + 0.316 MB written


def flops_per_net_pass(S, hidden, AD):
    dims = [S, hidden[0]] + list(hidden) + [AD]
    return 2 * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(hbm=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sust=p['bf16_tflops_sustained'], src='measured')
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src='fallback')


class ClockSampler:
    """nvidia-smi sampler running DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[2:6]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def make_oracle(cfg, n, seed):
    from oracle.sf_oracle import OracleSF
    gen = torch.Generator().manual_seed(seed)
    o = OracleSF(cfg['S'], cfg['A'], cfg['D'], cfg['hidden'], cfg['acts'], tsf_dim=cfg['gdim'], beta=cfg['beta'])
    for _ in range(n):
        o.add_random_policy(gen)
    return o, gen


def cpu_all_task_update(o, tr):
    """The all-task update the way the reference does it: one update_successor per task on the same batch."""
    for i in range(o.n_tasks):
        o.tsf_update_successor(tr, i, True)


def time_cpu(cfg, n_policies, steps, warmup, threads):
    from oracle.sf_oracle import synthetic_transitions
    torch.set_num_threads(threads)
    o, gen = make_oracle(cfg, n_policies, SEED)
    batches = [synthetic_transitions(cfg['B'], cfg['S'], cfg['A'], cfg['D'], gen) for _ in range(2)]
    for k in range(warmup):
        cpu_all_task_update(o, batches[k % 2])
    t0 = time.perf_counter()
    for k in range(steps):
        cpu_all_task_update(o, batches[k % 2])
    dt = time.perf_counter() - t0
    return cfg['B'] * n_policies * steps / dt, dt / steps


def run_reference(args, cfg, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_total = cfg['n_local'] * world
    steps = max(1, min(args.steps, 10))
    warm = max(1, min(args.warmup, 2))
    val, per = time_cpu(cfg, n_total, steps, warm, cores)
    sample = f'{steps} all-task steps ({n_total} update_successor calls each) of the same workload, B={cfg["B"]}'
    print(json.dumps({
        'impl': 'reference', 'metric': 'SF TD updates (transitions x tasks)/s', 'value': val, 'unit': 'updates/s',
        'n_gpus': world, 'steps': steps, 'warmup': warm, 'ms_per_step': per * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.workload, cfg, n_total, world),
        'cpu_baseline': {'value': val, 'unit': 'updates/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_config(name, cfg, n_total, world, precision='fp32', l2='flush', exchange=None):
    extra = {'exchange': exchange} if exchange else {}
    return {**extra, 'workload': f'{name}: TSFDQN Reacher S4/A9/D12 MLP 256-256 relu, g 4->100, h 100->12, beta={cfg["beta"]}, B={cfg["B"]}, '
                        f'{cfg["n_local"]} policies/GPU ({n_total} total), all-task fused TD update with GPI next actions',
            'batch': cfg['B'], 'policies_total': n_total, 'policies_per_gpu': cfg['n_local'],
            'parallelism': f'policy-sharded x{world}' if world > 1 else 'single GPU',
            'l2': ('flushed between timed steps (256 MiB write), flush excluded from the per-step CUDA-event time' if l2 == 'flush'
                   else 'inputs larger than L2: >160 MB of distinct resident batches cycled, weights / optimizer state stay '
                        'L2-resident as in a real training loop'),
            'precision_mode': ('bf16 operands on tcgen05, fp32 accumulate, fp32 master weights / backward / Adam (tolerance 2e-2)'
                               if precision == 'bf16' else 'fp32 (1e-5 parity mode)')}


def build_agent(cfg, n_local, precision='fp32'):
    from tests.gpu_util import FakeTask, model_lambda, HYPER
    from deep_successor_features_for_transfer_b200.tsfdqn import DeepTSF, TSFDQN, ReplayBuffer
    hyper = dict(HYPER, g_h_function_dims=cfg['gdim'], beta_loss_coefficient=cfg['beta'], precision=precision)
    dsf = DeepTSF(pytorch_model_handle=model_lambda(cfg['hidden'], cfg['acts']), use_true_reward=False,
                  target_update_ev=1000, hyperparameters=hyper)
    ag = TSFDQN(deep_sf=dsf, buffer_handle=lambda: ReplayBuffer(), gamma=0.9, T=500, encoding=None, use_gpi=True,
                hyperparameters=hyper)
    ag.reset()
    for i in range(n_local):
        ag.add_training_task(FakeTask(cfg['S'], cfg['A'], cfg['D'], i))
    return dsf, ag


GPI_EVAL = dict(S=11, A=27, D=50, hidden=[256, 256], acts=['relu', 'relu'], N=64, B=65536)


def time_gpi_eval(world, rank, dev, precision, barrier, reps=10):
    from tests.gpu_util import FakeTask, model_lambda, HYPER
    from deep_successor_features_for_transfer_b200.sfdqn import DeepSF
    import torch.distributed as dist
    c = GPI_EVAL
    if c['N'] % world:
        return {'skipped': f'{c["N"]} policies do not split over {world} ranks'}
    n_local = c['N'] // world
    torch.manual_seed(SEED + 100 + rank)
    sf = DeepSF(pytorch_model_handle=model_lambda(c['hidden'], c['acts']), hyperparameters=dict(HYPER, precision=precision))
    sf.reset()
    for i in range(n_local):
        sf.add_training_task(FakeTask(c['S'], c['A'], c['D'], rank * n_local + i))
    lib = sf._library
    if world > 1:
        lib.enable_sharding()
    gen = torch.Generator().manual_seed(SEED + 7)
    x = torch.sigmoid(torch.randn(c['B'], c['S'], generator=gen)).to(dev)          # tasks/hopper_phi.py:59
    w = torch.empty(c['D']).uniform_(-0.01, 0.01, generator=gen).to(dev)           # sfdqn.py:197
    for _ in range(3):
        lib.gpi(x, w, want_q=False)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        _, ka, kt = lib.gpi(x, w, want_q=False)
        b.record()
    barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    best_action = lib.decode_keys(ka)                                             # the result a caller reads
    F = flops_per_net_pass(c['S'], c['hidden'], c['A'] * c['D']) + 2 * c['A'] * c['D']      # the reference's arithmetic
    F_exec = flops_per_net_pass(c['S'], c['hidden'], (c['A'] + 15) // 16 * 16)               # with w folded into the output layer
    return {'metric': 'GPI action evals/s', 'value': c['B'] / (ms * 1e-3), 'unit': 'states/s (each over all 64 policies)',
            'ms_per_eval': ms, 'algorithmic_tflops_per_gpu': n_local * c['B'] * F / (ms * 1e-3) / 1e12,
            'executed_tflops_per_gpu': n_local * c['B'] * F_exec / (ms * 1e-3) / 1e12, 'scaling': 'strong',
            'note': 'algorithmic = N*F per state as the reference computes it (psi[B,N,A,D] then .w); executed = what the kernel '
                    'runs after folding w into the output layer (A columns instead of A*D), so algorithmic can exceed the bf16 peak',
            'config': f'SFDQN Hopper S11/A27/D50, {c["N"]} policies ({n_local}/GPU), B={c["B"]} resident states, pack + fold + fused '
                      f'forward/GPI kernel + key all-reduce per eval, {precision}',
            'check': int(best_action.min()) >= 0 and int(best_action.max()) < c['A']}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='tsfdqn_reacher_b4096')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpi-eval', action='store_true', help='skip the secondary M2 measurement (quick A/B runs)')
    ap.add_argument('--l2', default='flush', choices=['flush', 'inputs'],
                    help='flush: 256 MiB write between timed steps (everything cold, weights included); inputs: cycle through '
                         'more distinct resident batches than fit in L2 (inputs cold, weights stay L2-resident)')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help='bf16: tcgen05 tensor-core forwards (stated tolerance 2e-2); fp32: CUDA-core 1e-5 parity mode')
    args = ap.parse_args()
    cfg = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    strong = 'n_total' in cfg
    if strong:
        if cfg['n_total'] % world:
            raise SystemExit(f'{cfg["n_total"]} policies do not split over {world} ranks')
        cfg['n_local'] = cfg['n_total'] // world
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, cfg, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    import torch.distributed as dist
    from tests.synthetic import synthetic_transitions
    from deep_successor_features_for_transfer_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(SEED + rank)
    B, S, A, D, n_local = cfg['B'], cfg['S'], cfg['A'], cfg['D'], cfg['n_local']
    n_total = n_local * world
    dsf, ag = build_agent(cfg, n_local, args.precision)
    lib = dsf._library
    if world > 1:
        lib.enable_sharding()
    gen = torch.Generator().manual_seed(SEED)                       # identical batches on every rank (replicated replay)
    host = [synthetic_transitions(B, S, A, D, gen) for _ in range(8)]
    pinned = [tuple(t.pin_memory() for t in tr) for tr in host]
    batch_bytes = sum(t.numel() * t.element_size() for t in host[0])
    n_res = 8 if args.l2 == 'flush' else int(160e6 // batch_bytes) + 1        # 'inputs': > 126 MB L2 worth of batches
    resident = [tuple(t.to(dev) for t in host[k % 8]) for k in range(n_res)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: HBM-resident inputs, CUDA events per step ----------------
    # nvidia-smi needs ~0.1 s to start and the timed regions last milliseconds: it samples every 20 ms from here (before the
    # warm-up) to the end of the e2e loop, i.e. throughout both timed regions and the loaded phases around them.
    sampler = ClockSampler(local_rank)
    sampler.start()
    n_warm = max(args.warmup, 1000)                     # a FIXED count (identical on every rank): >= W steps, ~0.3 s under load
    for k in range(n_warm):
        ag.update_successor_all(resident[k % n_res], use_gpi=True)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # probes around the dominant kernel (the one-launch fused forward: online psi(s) + GPI(s') + target psi(s')), recorded by
    # the command list itself inside every timed step
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    plan_key = lib.last_plan_key
    l0 = _lib.launch_count
    barrier()
    for k in range(args.steps):
        if args.l2 == 'flush':
            flush.fill_(k & 0xFF)
        lib.set_probe(plan_key, *kev[k])
        ev[k][0].record()
        ag.update_successor_all(resident[(n_warm + k) % n_res], use_gpi=True)
        ev[k][1].record()
    barrier()
    lib.set_probe(plan_key, None, None)
    launches = _lib.launch_count - l0
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax)
    updates_per_step = B * n_total
    value = updates_per_step * args.steps / (total_ms * 1e-3)

    # ---------------- e2e: public API from pinned host batches, losses read back every step ----------------
    e2e_steps = args.steps
    losses_host = torch.zeros(n_local, 3).pin_memory()              # the step's command list ends with the D2H copy into it
    stream = torch.cuda.current_stream()
    for k in range(3):
        ag.update_successor_all(pinned[k % 8], use_gpi=True, host_losses=losses_host)
        stream.synchronize()
    barrier()
    check = 0.0
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        ag.update_successor_all(pinned[k % 8], use_gpi=True, host_losses=losses_host)
        stream.synchronize()                                        # the step's result is on the host from here on
        check += float(losses_host[0, 0])                           # (read it: D2H of the losses every step)
    barrier()
    assert check == check and check > 0.0, 'e2e losses did not reach the host'
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = updates_per_step * e2e_steps / float(e2e_t)
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])
    d2h = losses_host.numel() * 4

    # ---------------- roofline of the dominant kernel ----------------
    # bf16 mode: mlp_forward_tc_kernel, ONE launch per step carrying 3 of the step's 5 net passes per (transition, policy):
    # online psi(s), GPI over the library on s' under every task's reward vector, target psi(s').  fp32 mode: the GPI forward.
    # Algorithmic FLOPs: F per (state, policy) pass + 2*A*D per (state, policy, reward vector) for psi . w.
    F = flops_per_net_pass(S, cfg['hidden'], A * D)
    peaks = load_peaks()
    if args.precision == 'bf16':
        k_flops = n_local * B * (3 * F + 2 * A * D * n_total)
        kname = ('mlp_forward_tc_kernel: one launch = online psi(s) + fused GPI(s\') + target psi(s\') for all local policies '
                 '(tcgen05 bf16 MMA, TMEM accumulators, TMA weight ring)')
    else:
        k_flops = n_local * B * (F + 2 * A * D * n_total)
        kname = 'mlp_forward_kernel (GPI form: fused ensemble MLP + GPI epilogue, fp32 CUDA-core mode)'
    achieved = k_flops / (k_ms * 1e-3) / 1e12
    roofline = {'kernel': kname, 'bound': 'tensor',
                'achieved': achieved, 'peak': peaks['tf_burst'], 'unit': 'TFLOP/s', 'frac': achieved / peaks['tf_burst'],
                'traffic': KERNEL_DRAM_BYTES.get(args.precision), 'peak_source': f'{peaks["src"]} bf16 cuBLAS burst',
                'kernel_ms': k_ms, 'flops_per_launch': k_flops,
                'timing': 'CUDA events recorded by the command list around the kernel inside every timed step (mean)',
                'note': 'latency-bound at this size: 3 x 128 row tiles on 148 SMs, ~2.6 tiles per SM, each a 4-layer dependent '
                        'chain; the tensor-pipe roofline is approached only at N >= 32 policies (scripts/tc_probe.py, DESIGN.md)'}
    step_flops = 5 * F * B * n_local                               # N GPI + N online + N target + 2N backward = 5N passes
    step_tflops = step_flops * args.steps / (total_ms * 1e-3) / 1e12

    # ---------------- M2: GPI action evals/s (BASELINE config 3) ----------------
    # Hopper shapes (S=11, A=27, D=50), 64 source policies sharded over the ranks (strong scaling: 64 / world per GPU), 65 536
    # states = sigmoid(N(0,1)) resident in HBM; one eval = fused ensemble forward + GPI epilogue for every state over ALL 64
    # policies, packed (value, index) keys MAX-all-reduced across ranks (NCCL).  CUDA events, max over ranks.
    gpi_eval = None
    try:
        gpi_eval = {'skipped': '--no-gpi-eval'} if args.no_gpi_eval else time_gpi_eval(world, rank, dev, args.precision, barrier)
    except Exception as e:                                          # the headline line must not depend on the secondary metric
        gpi_eval = {'error': f'{type(e).__name__}: {e}'}

    out = None
    if rank == 0:
        out = {
            'metric': 'SF TD updates (transitions x tasks)/s', 'value': value, 'unit': 'updates/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if strong else 'weak', 'vs_baseline': None, 'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
            'config': workload_config(args.workload, cfg, n_total, world, args.precision, args.l2,
                                      None if world == 1 else ('peer-memory kernels in the step\'s launch chain (CUDA IPC arenas, signal/wait '
                                      'flags, 128-bit pulls over NVLink): GPI keys MAX reduce-scatter + [w | delta h] all-gather, no NCCL call '
                                      'inside a step' if lib._peer is not None else 'NCCL: all-reduce(MAX) of packed keys + one all-gather')),
            'clocks': clocks,
            'e2e': {'value': e2e_val, 'unit': 'updates/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': float(e2e_t) / e2e_steps * 1e3},
            'warmup_steps_run': n_warm,                     # >= --warmup: a fixed count on every rank, clocks settled under load
            'gpu_launches': launches,
            'roofline': roofline,
            'step_tflops_per_gpu': step_tflops,
            'gpi_eval': gpi_eval,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cval, cper = time_cpu(cfg, n_total, 10 if n_total <= 8 else 1, 2 if n_total <= 8 else 0, cores)
            try:                                                    # "also a 1-thread number" (3 steps)
                c1 = time_cpu(cfg, n_total, 3, 1, 1)[0] if n_total <= 8 else None      # (large ensembles: minutes at one thread)
            except Exception:
                c1 = None
            out['cpu_baseline'] = {'value': cval, 'unit': 'updates/s', 'cores': cores, 'kind': 'port', 'value_1_thread': c1,
                                   'sample': f'10 all-task steps ({n_total} update_successor calls each), B={B}, '
                                             f'{cper * 1e3:.1f} ms/step, oracle port on torch CPU fp32'}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
