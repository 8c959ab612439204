"""SM clock seen by the forward kernel (clock64 / globaltimer) in isolated steps and in a back-to-back stream of steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deep_successor_features_for_transfer_b200.workloads import build_tsf_agent, synthetic_transitions
dsf, ag = build_tsf_agent('reacher', 4, precision='bf16', seed=7)
gen = torch.Generator().manual_seed(1)
tr = tuple(t.cuda() for t in synthetic_transitions(4096, 4, 9, 12, gen))
os.environ.pop('X', None)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for k in range(n):
    ag.update_successor_all(tr, use_gpi=True)
torch.cuda.synchronize()
