"""Per-layer error report of the bf16 tcgen05 CNN path against torch fp32 (debug aid, needs a GPU)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import GraspEngine, pack_weights
from test_gpu_parity import _torch_layer_features

n = int(sys.argv[1]) if len(sys.argv) > 1 else 61
sd = O.seeded_state_dict(1234)
rng = np.random.default_rng(5)
x = torch.from_numpy(rng.random((n, 9, 32, 32), dtype=np.float32))
x[:, 1] = (x[:, 1] > 0.5).float()
eng = GraspEngine(1, 64, 64, 2)
eng.set_cnn_weights(pack_weights(sd))
for layer in range(6):
    got = eng.cnn_bf16_features(x, layer).cpu()
    torch.cuda.synchronize()
    with torch.no_grad():
        want = _torch_layer_features(sd, x, layer)
    if layer == 5:
        want = want.permute(0, 2, 3, 1).contiguous()
    d = (got - want).abs()
    print(f"layer {layer}: shape {tuple(got.shape)} scale {float(want.abs().max()):.4f} max_err {float(d.max()):.5f} "
          f"mean_err {float(d.mean()):.6f} nonzero got {float((got != 0).float().mean()):.3f} want {float((want != 0).float().mean()):.3f}", flush=True)
    if float(d.max()) > 0.05 * float(want.abs().max()):
        bad = torch.nonzero(d > 0.05 * float(want.abs().max()))
        print("  first bad idx", bad[:8].tolist(), "n_bad", len(bad))
        print("  got", got[tuple(bad[0])].item(), "want", want[tuple(bad[0])].item())
y16 = eng.cnn_forward(x, use_bf16=True).cpu().numpy()
y32 = eng.cnn_forward(x, use_bf16=False).cpu().numpy()
print("logits max abs diff bf16 vs fp32:", float(np.abs(y16 - y32).max()), "logit scale", float(np.abs(y32).max()))
for m in (5120, 65536 // 8):
    xb = torch.from_numpy(rng.random((m, 9, 32, 32), dtype=np.float32)).cuda()
    for use in (True, False):
        eng.cnn_forward(xb, use_bf16=use); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.cnn_forward(xb, use_bf16=use); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"n={m} bf16={use}: {ms:.3f} ms  {m * 312.83e6 / ms / 1e9:.1f} TFLOP/s")
