"""A/B timing of the tensor-core CNN (LG_CNN_ISSUERS=1|2|4 selects the number of MMA-issuing warps): python tools/cnn_ab.py [patches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import GraspEngine, pack_weights
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3726
eng = GraspEngine(256, 64, 64, 2)
eng.set_cnn_weights(pack_weights(O.seeded_state_dict(1234)))
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.rand((n, 9, 32, 32), device="cuda", generator=g)
ref = eng.cnn_forward(x[:512], use_bf16=False)
for _ in range(3):
    y = eng.cnn_forward(x, use_bf16=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(10):
    e0.record(); y = eng.cnn_forward(x, use_bf16=True); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print(f"issuers={os.environ.get("LG_CNN_ISSUERS","4")} n={n}: {ms:.3f} ms  {312.83e6*n/ms/1e9:.0f} TFLOP/s  max|bf16-fp32|={float((y[:512]-ref).abs().max()):.4f}", flush=True)
