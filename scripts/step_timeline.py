"""In-kernel clock64 timeline (CTA 0) of the forward kernel inside a headline train step: SFGPI_TIMELINE=1 python scripts/step_timeline.py [N] [B] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_successor_features_for_transfer_b200.workloads import build_tsf_agent, synthetic_transitions

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = sys.argv[4] if len(sys.argv) > 4 else 'bf16'
dsf, ag = build_tsf_agent('reacher', N, precision=prec, seed=7)
gen = torch.Generator().manual_seed(1)
tr = tuple(t.cuda() for t in synthetic_transitions(B, 4, 9, 12, gen))
for k in range(steps):
    print(f'--- step {k}', file=sys.stderr, flush=True)
    ag.update_successor_all(tr, use_gpi=True)
    torch.cuda.synchronize()

# SFGPI_TRACE=1: kernel windows of the last steps (globaltimer), see csrc/common.cuh
if os.environ.get('SFGPI_TRACE'):
    from deep_successor_features_for_transfer_b200 import _lib
    fn = _lib.lib().sfgpi_trace_dump
    fn.restype = None
    fn()
    for k in range(3):
        ag.update_successor_all(tr, use_gpi=True)
        fn()
    # back-to-back steps (the device queue stays full): windows of 20 steps merged = min entry of the first, max exit of the last
    for k in range(20):
        ag.update_successor_all(tr, use_gpi=True)
    fn()
