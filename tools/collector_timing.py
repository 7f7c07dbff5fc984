"""Cost of lg_collect_samples beside lg_process_batch on the metric's workload (256 frames 1440x1080, 30 leaves), and a
full-size parity check of two frames against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, torch
import leafgrasp_oracle as O
from leafgrasp_b200 import GraspEngine, camera_from_projection, pack_weights, synth, _native as N

spec, B, U = synth.CFG2, 256, 32
P = synth.projection_matrix(spec)
cam = camera_from_projection(P)
lab, dep = synth.make_batch(spec, 7, 0, U)
lab = np.concatenate([lab] * (B // U)); dep = np.concatenate([dep] * (B // U))
lt, dt = torch.from_numpy(lab).cuda(), torch.from_numpy(dep).cuda()
eng = GraspEngine(B, spec.height, spec.width, 128)
eng.set_cnn_weights(pack_weights(O.seeded_state_dict(1234)))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tp, tc = [], []
for it in range(6):
    ev[0].record()
    res = eng.process_batch(lt, dt, cam, True, sync=False)
    ev[1].record()
    patches, meta, sizes = eng.collect_samples(dt, labels=lt, seed=1, first_frame_index=it * B)
    ev[2].record()
    torch.cuda.synchronize()
    if it >= 2:
        tp.append(ev[0].elapsed_time(ev[1])); tc.append(ev[1].elapsed_time(ev[2]))
nvalid = int(meta["valid"].sum())
print(f"process_batch {np.mean(tp):.2f} ms, collect_samples {np.mean(tc):.2f} ms for {B} frames "
      f"({nvalid} samples, {nvalid / (np.mean(tc) * 1e-3):.0f} samples/s); set sizes mean {sizes.mean(axis=0).round(1).tolist()} "
      f"max {sizes.max(axis=0).tolist()}", flush=True)
# parity at full size: frames 0 and 1
res = eng._records(res, N.FRAME_RESULT)
pt = patches.cpu().numpy()
f, cx, cy = P[0, 0], P[0, 2], P[1, 2]
for b in (0, 1):
    mask = (lab[b] == res["leaf_id"][b]).astype(np.uint8)
    s = O.score_maps(mask, dep[b], f, cx, cy, "strict")
    g = (int(res["grasp_x"][b]), int(res["grasp_y"][b]))
    want = O.collect_sample(mask, dep[b], s, g, float(np.max(s["traditional_score"])), O.CollectorRng(1, 5 * B + b))
    valid = [k for k in range(7) if meta[b, k]["valid"]]
    assert len(valid) == len(want)
    for k, w in zip(valid, want):
        assert (int(meta[b, k]["x"]), int(meta[b, k]["y"])) == tuple(w["grasp_point"])
        for ch in (0, 1, 5, 6, 8):
            if not (ch == 0 and w["is_augmented"]):
                assert np.array_equal(pt[b, k, ch], w["patch"][ch])
        np.testing.assert_allclose(pt[b, k], w["patch"], rtol=1e-5, atol=1e-6)
    print(f"frame {b}: {len(want)} samples identical to the oracle; sets {sizes[b].tolist()}", flush=True)
