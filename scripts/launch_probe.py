"""Back-to-back launches of one tensor-core forward (psi form, no fold kernel in between): per-launch time vs launch count."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as gu
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.library import _stream, ptr
from deep_successor_features_for_transfer_b200.sfdqn import DeepSF

S, A, D, N, B = 4, 9, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 4096
sf = DeepSF(pytorch_model_handle=gu.model_lambda([256, 256], ['relu', 'relu']), hyperparameters=dict(gu.HYPER, precision='bf16'))
sf.reset()
for i in range(N):
    sf.add_training_task(gu.FakeTask(S, A, D, i))
lib = sf._library
x = torch.randn(B, S, device='cuda')
acts = torch.randint(0, A, (B,), device='cuda')
sel = torch.empty(N, B, D, device='cuda')
a = lib._fwd_args(lib.online, 0, N, x)
a.sel_actions, a.sel_out = acts.data_ptr(), sel.data_ptr()
lib._pack('online', 0, N)
call = lambda: _lib.call('sfgpi_mlp_forward_tc', C.byref(a), ptr(lib._shadow_for('online')), lib.cap, None, None, _stream())
small = torch.empty(1024, device='cuda')
for _ in range(5):
    call()
torch.cuda.synchronize()
for reps, interleave in ((1, False), (4, False), (16, False), (16, True)):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(5):
        s.record()
        for _ in range(reps):
            call()
            if interleave:
                small.fill_(1.0)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3 / reps)
    print(f'N={N} B={B} reps={reps:2d} interleave_small_kernel={interleave}: {sorted(ts)[2]:.1f} us per forward launch', flush=True)
