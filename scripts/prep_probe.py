"""Times the pieces of the step prologue at BASELINE config 4 sizes (N policies, n_w reward vectors): GPI fold, bf16 pack, key fill."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests import gpu_util as gu
from deep_successor_features_for_transfer_b200 import _lib
from deep_successor_features_for_transfer_b200.library import _stream
from deep_successor_features_for_transfer_b200.sfdqn import DeepSF

S, A, D = 4, 9, 12
N, B, NW = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sf = DeepSF(pytorch_model_handle=gu.model_lambda([256, 256], ['relu', 'relu']), hyperparameters=dict(gu.HYPER, precision='bf16'))
sf.reset()
for i in range(N):
    sf.add_training_task(gu.FakeTask(S, A, D, i))
lib = sf._library
desc = lib.spec.desc()
w = (torch.rand(NW, D, device='cuda') * 0.02 - 0.01).contiguous()
nq = _lib.lib().sfgpi_gpi_fold_rows(C.byref(desc), NW)
wq = torch.empty(N * nq * 256, dtype=torch.bfloat16, device='cuda')
bq = torch.empty(N * nq, device='cuda')
keys = torch.empty(NW, B, dtype=torch.int64, device='cuda')
shadow = lib._shadow_for('online')


def timeit(name, fn, nbytes):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f'{name:10s} {t * 1e3:9.1f} us   {nbytes / 1e6:8.1f} MB   {nbytes / t / 1e6:8.1f} GB/s', flush=True)


timeit('fold', lambda: _lib.call('sfgpi_fold_gpi', C.byref(desc), lib.online.data_ptr(), 0, N, w.data_ptr(), NW, 0, wq.data_ptr(), bq.data_ptr(), _stream()),
       N * (4 * A * D * 256 + 2 * 256 * nq))
timeit('pack', lambda: _lib.call('sfgpi_pack_bf16', C.byref(desc), lib.online.data_ptr(), 0, N, shadow.data_ptr(), _stream()), N * 6 * lib.spec.n_params)
timeit('keys_fill', lambda: _lib.call('sfgpi_keys_fill', keys.data_ptr(), keys.numel(), _stream()), keys.numel() * 8)
