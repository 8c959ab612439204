"""Where the end-to-end loop's time goes beyond the device step: variants of the pinned-batch loop (headline workload)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_successor_features_for_transfer_b200.workloads import build_tsf_agent, synthetic_transitions

N, B, K = 4, 4096, 400
dsf, ag = build_tsf_agent('reacher', N, precision='bf16', seed=7)
lib = dsf._library
gen = torch.Generator().manual_seed(1)
dev = [tuple(t.cuda() for t in synthetic_transitions(B, 4, 9, 12, gen)) for _ in range(8)]
pin = [tuple(t.cpu().pin_memory() for t in tr) for tr in dev]
lh = [torch.zeros(N, 3).pin_memory() for _ in range(2)]
stream = torch.cuda.current_stream()
for k in range(300):
    ag.update_successor_all(dev[k % 8], use_gpi=True)
torch.cuda.synchronize()


def timed(name, body):
    body(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    body(K)
    torch.cuda.synchronize()
    print(f'{name:60s} {(time.perf_counter() - t0) / K * 1e6:8.1f} us / step', flush=True)


def loop(src, host_losses, events):
    def body(n):
        evs = [torch.cuda.Event() for _ in range(2)]
        for k in range(n):
            ag.update_successor_all(src[k % 8], use_gpi=True, host_losses=lh[k & 1] if host_losses else None)
            if events:
                evs[k & 1].record(stream)
                if k > 0:
                    evs[(k - 1) & 1].synchronize()
    return body


# measured on B200 (round 2): 100.8 / 103.7 / 103.2 / 109.9 / 109.8 us per step: the losses' zero-copy store costs ~3 us, the
# in-prologue PCIe pull of the 393 KB batch ~6 us, the per-step event + lagged wait nothing.  (Staging the batch one step ahead
# by a copy-only launch on a side stream was tried: bit-identical results, no gain -- 109.8 us -- and dropped.)
timed('device batches, no losses to host, no events', loop(dev, False, False))
timed('device batches, losses to pinned host, no events', loop(dev, True, False))
timed('device batches, losses to host, event per step + lagged wait', loop(dev, True, True))
timed('pinned batches (in-prologue pull), losses to host, no events', loop(pin, True, False))
timed('pinned batches (in-prologue pull), events + lagged wait', loop(pin, True, True))
