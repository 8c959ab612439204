"""Times the tensor-core GPI forward alone as a function of the number of reward vectors n_w (N policies, B states)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as gu
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.library import _stream
from deep_successor_features_for_transfer_b200.sfdqn import DeepSF

S, A, D = 4, 9, 12
N, B = int(sys.argv[1]), int(sys.argv[2])
sf = DeepSF(pytorch_model_handle=gu.model_lambda([256, 256], ['relu', 'relu']), hyperparameters=dict(gu.HYPER, precision='bf16'))
sf.reset()
for i in range(N):
    sf.add_training_task(gu.FakeTask(S, A, D, i))
lib = sf._library
x = torch.randn(B, S, device='cuda')
lib._pack('online', 0, N)
for nw in [int(v) for v in sys.argv[3:]]:
    w = (torch.rand(nw, D, device='cuda') * 0.02 - 0.01).contiguous()
    keys = torch.empty(nw, B, dtype=torch.int64, device='cuda')
    a = lib._fwd_args(lib.online, 0, N, x)
    a.w, a.n_w, a.w_diag = w.data_ptr(), nw, 0
    a.key_action = keys.data_ptr()
    wq, bq = lib._fold(a, 'online')
    torch.cuda.synchronize()
    ts = []
    for rep in range(4):
        _lib.call('sfgpi_keys_fill', keys.data_ptr(), keys.numel(), _stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call('sfgpi_mlp_forward_tc', C.byref(a), lib._shadow_for('online').data_ptr(), lib.cap, wq.data_ptr(), bq.data_ptr(), _stream())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    tiles = N * ((B + 127) // 128)
    print(f'N={N} B={B} n_w={nw:4d}: {min(ts):8.3f} ms  ({min(ts) * 1e-3 * 1.965e9 * 148 / tiles / 1e3:7.1f} k SM-cycles per tile)', flush=True)
