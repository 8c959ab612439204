"""Times the fused GPI forward (fp32 CUDA-core vs tcgen05 bf16) at a few ensemble sizes; prints TFLOP/s of algorithmic FLOPs."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as gu
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.library import _stream

def run(S, A, D, N, B, precision, reps=10):
    meta = dict(S=S, A=A, D=D, hidden=[256, 256], acts=['relu', 'relu'], N=N)
    from deep_successor_features_for_transfer_b200.sfdqn import DeepSF
    sf = DeepSF(pytorch_model_handle=gu.model_lambda(meta['hidden'], meta['acts']), hyperparameters=dict(gu.HYPER, precision=precision))
    sf.reset()
    for i in range(N):
        sf.add_training_task(gu.FakeTask(S, A, D, i))
    lib = sf._library
    x = torch.randn(B, S, device='cuda')
    keys = torch.empty(2, B, dtype=torch.int64, device='cuda')
    a = lib._fwd_args(lib.online, 0, N, x)
    a.w, a.n_w, a.w_diag = lib.w[0].data_ptr(), 1, 0
    a.key_action, a.key_task = keys[0].data_ptr(), keys[1].data_ptr()
    if precision == 'bf16':
        lib._pack('online', 0, N)
    for _ in range(3):
        lib._forward(a, 'online', fresh=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for s, e in ev:
        _lib.call('sfgpi_keys_fill', keys.data_ptr(), 2 * B, _stream())
        s.record()
        lib._forward(a, 'online', fresh=True)
        e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in ev)[reps // 2]
    F = 2 * (S * 256 + 2 * 256 * 256 + 256 * A * D) + 2 * A * D
    tf = N * B * F / (ms * 1e-3) / 1e12
    print(f'{precision:5s} S={S} A={A} D={D} N={N:3d} B={B:6d}: {ms:9.3f} ms  {tf:8.1f} TFLOP/s  ({B / (ms * 1e-3) / 1e6:8.2f} M GPI evals/s)', flush=True)

if __name__ == '__main__':
    if len(sys.argv) > 1:
        S, A, D, N, B = (int(v) for v in sys.argv[2:7])
        run(S, A, D, N, B, sys.argv[1], reps=3)
        sys.exit(0)
    for prec in ('bf16', 'fp32'):
        run(4, 9, 12, 4, 4096, prec)
        run(4, 9, 12, 32, 4096, prec)
        run(4, 9, 12, 64, 16384, prec, reps=5)
        run(11, 27, 50, 8, 65536, prec, reps=3)
