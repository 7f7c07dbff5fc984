"""Where the tcgen05 conv kernels spend their cycles.  Needs liblgb200_timing.so (`make -C leaf-grasping-vision-ml_b200/csrc
timing`: the conv kernels with clock64 timers around every mbarrier wait).  Prints, per layer, the share of the MMA
warp's time spent waiting for the accumulators / input stage / weight stages and issuing, and the epilogue's wait / work."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import _native as N
N.LIB_PATH = os.path.join(ROOT, "leaf-grasping-vision-ml_b200", "liblgb200_timing.so")
from leafgrasp_b200 import GraspEngine, pack_weights

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3726
eng = GraspEngine(256, 64, 64, 2)
eng.set_cnn_weights(pack_weights(O.seeded_state_dict(1234)))
x = torch.rand((n, 9, 32, 32), device="cuda")
for _ in range(3):
    eng.cnn_forward(x, use_bf16=True)
torch.cuda.synchronize()
buf = np.zeros((6, 148, 8), dtype=np.uint64)
lib = N.lib()
lib.lg_cnn_timing.argtypes = [C.c_void_p]
assert lib.lg_cnn_timing(buf.ctypes.data_as(C.c_void_p)) == 0
names = ["wait acc_empty", "wait a_full", "wait b_full", "issue", "epi wait acc_full", "epi work", "mma warp total"]
for l in range(6):
    t = buf[l].astype(np.float64)
    act = t[:, 6] > 0
    tot = t[act, 6].mean()
    print(f"layer {l}: total {tot:9.0f} cyc | " + " | ".join(f"{names[k]} {t[act, k].mean() / tot * 100:5.1f}%" for k in range(6)))
