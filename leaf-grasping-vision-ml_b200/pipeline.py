"""GraspEngine: one device context of the native library + thin tensor-in / record-out calls.

torch is used for device memory, streams and (in dist.py) the NCCL plumbing only; every computation
happens inside liblgb200.so.  All calls run on the current torch CUDA stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def camera_from_projection(P) -> N.Camera:
    """f = P[0,0], cx = P[0,2], cy = P[1,2] (grasp_point_selector.py:145-150)."""
    P = np.asarray(P, dtype=np.float64)
    return N.Camera(float(P[0, 0]), float(P[0, 2]), float(P[1, 2]))


class GraspEngine:
    """lanes > 1: process_batch splits a batch into `lanes` contiguous parts, part 0 on the caller's stream and the main
    context, the others on their own stream and context.  Frames are independent, so the parts only meet in the result
    buffer; running them side by side lets one part's latency-bound stages (per-frame sweeps, the 20 arg-max rounds)
    fill the gaps of the other's.  Every other call uses the main context."""

    def __init__(self, max_frames: int, height: int, width: int, max_labels: int = 128, device=None, lanes: int = 1):
        if not torch.cuda.is_available():
            raise N.NativeError("GraspEngine needs a CUDA device; this package has no CPU fallback")
        self.lib = N.lib()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.B, self.H, self.W, self.L = int(max_frames), int(height), int(width), int(max_labels)
        self._ctx = C.c_void_p(0)
        self._lane_ctx, self._lane_streams = [], []
        self.lanes_active = True
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_create(C.byref(self._ctx), self.B, self.H, self.W, self.L), "lg_create")
            per = -(-self.B // max(1, int(lanes)))
            for _ in range(1, max(1, int(lanes))):
                ctx = C.c_void_p(0)
                N.check(self.lib.lg_create(C.byref(ctx), per, self.H, self.W, self.L), "lg_create")
                self._lane_ctx.append(ctx)
                self._lane_streams.append(torch.cuda.Stream(device=self.device))
        self.has_cnn = False
        self._rec_out = None

    def _all_ctx(self):
        return [self._ctx] + self._lane_ctx

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            with torch.cuda.device(self.device):
                torch.cuda.synchronize()
                for ctx in self._all_ctx():
                    self.lib.lg_destroy(ctx)
            self._ctx = C.c_void_p(0)
            self._lane_ctx = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_overlap(self, on: bool):
        """Run independent stages side by side on the library's internal stream (default) or serialised."""
        for ctx in self._all_ctx():
            N.check(self.lib.lg_set_overlap(ctx, int(bool(on))), "lg_set_overlap")

    CODE_WEIGHTS = dict(approach=0.4, sdf=0.3, flatness=0.2, accessibility=0.1)       # grasp_point_selector.py:272-277
    README_WEIGHTS = dict(approach=0.40, sdf=0.20, flatness=0.25, accessibility=0.15)   # reference README.md:83-87

    def set_score_weights(self, approach=0.4, sdf=0.3, flatness=0.2, accessibility=0.1):
        """Weights of the traditional score.  Default: the reference's code; GraspEngine.README_WEIGHTS is the set its
        README advertises (SURVEY.md 8a': selectable, not parity-tested against the reference - its code never uses it)."""
        for ctx in self._all_ctx():
            N.check(self.lib.lg_set_score_weights(ctx, float(approach), float(sdf), float(flatness), float(accessibility)),
                    "lg_set_score_weights")

    def set_host_label_rle(self, on: bool):
        """process_batch_host: run-length encode the label images on the host inside the call (default) or copy them raw."""
        for ctx in self._all_ctx():
            N.check(self.lib.lg_set_host_label_rle(ctx, int(bool(on))), "lg_set_host_label_rle")

    def host_call_bytes(self):
        """(host-to-device, device-to-host) bytes of the last process_batch_host call."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        N.check(self.lib.lg_host_call_bytes(self._ctx, C.byref(a), C.byref(b)), "lg_host_call_bytes")
        return int(a.value), int(b.value)

    def set_patch_export(self, on: bool):
        """Drop-in mode (default) keeps the float32 patch tensor of the last call for last_patches(); throughput mode
        (False) lets the gather kernel write the tensor-core CNN's input directly (bf16 CNN only; same results)."""
        for ctx in self._all_ctx():
            N.check(self.lib.lg_set_patch_export(ctx, int(bool(on))), "lg_set_patch_export")

    @property
    def context_bytes(self) -> int:
        return sum(int(self.lib.lg_context_bytes(ctx)) for ctx in self._all_ctx())

    # ---- helpers ---------------------------------------------------------------------------------
    def _frames(self, t, dtype, name):
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.shape[-2:] != (self.H, self.W):
            raise ValueError(f"{name}: expected [..., {self.H}, {self.W}], got {tuple(t.shape)}")
        return t.to(self.device, dtype, non_blocking=True).contiguous()

    def _new_results(self, n):
        return torch.empty(n * N.FRAME_RESULT.itemsize, dtype=torch.uint8, device=self.device)

    @staticmethod
    def _records(buf, dtype):
        return np.frombuffer(buf.cpu().numpy().tobytes(), dtype=dtype)

    # ---- CNN ----------------------------------------------------------------------------------------
    def set_cnn_weights(self, blob, config: "N.CnnConfig | None" = None):
        """Folded weights (cnn.pack_weights) of the default architecture, or of the one `config` (N.cnn_config) names."""
        with torch.cuda.device(self.device):
            if blob is None:
                for ctx in self._all_ctx():
                    N.check(self.lib.lg_set_cnn_weights(ctx, None, 0), "lg_set_cnn_weights")
                self.has_cnn = False
                return
            blob = np.ascontiguousarray(blob, dtype=np.float32)
            for ctx in self._all_ctx():
                if config is None:
                    N.check(self.lib.lg_set_cnn_weights(ctx, blob.ctypes.data_as(C.c_void_p), blob.size), "lg_set_cnn_weights")
                else:
                    N.check(self.lib.lg_set_cnn_model(ctx, C.byref(config), blob.ctypes.data_as(C.c_void_p), blob.size),
                            "lg_set_cnn_model")
            self.has_cnn = True

    def cnn_forward(self, patches: torch.Tensor, use_bf16: bool = False) -> torch.Tensor:
        patches = patches.to(self.device, torch.float32).contiguous()
        n = patches.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_cnn_forward(self._ctx, _ptr(patches), n, _ptr(out), int(use_bf16), _stream()), "lg_cnn_forward")
        return out

    def cnn_bf16_features(self, patches: torch.Tensor, layer: int) -> torch.Tensor:
        """Activations of the bf16 tensor-core path after conv layer `layer` (see include/leafgrasp.h)."""
        patches = patches.to(self.device, torch.float32).contiguous()
        n = patches.shape[0]
        shape = [(64, 32, 32), (64, 16, 16), (128, 16, 16), (128, 8, 8), (256, 8, 8), (4, 4, 256)][layer]
        out = torch.empty((n,) + shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_cnn_bf16_features(self._ctx, _ptr(patches), n, layer, _ptr(out), _stream()),
                    "lg_cnn_bf16_features")
        return out

    def normalize_patches(self, raw: torch.Tensor) -> torch.Tensor:
        raw = raw.to(self.device, torch.float32).contiguous()
        out = torch.empty_like(raw)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_normalize_patches(self._ctx, _ptr(raw), raw.shape[0], _ptr(out), _stream()), "lg_normalize_patches")
        return out

    # ---- whole path ----------------------------------------------------------------------------------
    def process_batch(self, labels, depth, cam: N.Camera, use_bf16: bool = False, sync: bool = True):
        """labels int16 / depth float32 device tensors [n,H,W] -> structured ndarray (or the raw device
        buffer when sync=False)."""
        labels = self._frames(labels, torch.int16, "labels")
        depth = self._frames(depth, torch.float32, "depth")
        n = labels.shape[0]
        res = self._new_results(n)
        parts = self.lane_split(n)
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream()
            for k in range(1, len(parts)):           # before anything of this call is queued on the caller's stream:
                self._lane_streams[k - 1].wait_stream(cur)   # the lanes only wait for what precedes the call (the inputs)
            for k, (lo, hi) in enumerate(parts):
                ctx = self._ctx if k == 0 else self._lane_ctx[k - 1]
                stream = cur if k == 0 else self._lane_streams[k - 1]
                rp = C.c_void_p(res.data_ptr() + lo * N.FRAME_RESULT.itemsize)
                if self._lane_ctx and self._rec_out is not None:      # every part writes at its own frame offset
                    if self._rec_out.shape[0] < n:
                        raise ValueError("record buffer holds fewer frames than the batch")
                    N.check(self.lib.lg_set_record_output(ctx, C.c_void_p(self._rec_out.data_ptr() + lo * N.TOP_K * 16)),
                            "lg_set_record_output")
                N.check(self.lib.lg_process_batch(ctx, _ptr(labels[lo:hi]), _ptr(depth[lo:hi]), hi - lo, C.byref(cam), rp,
                                                  int(use_bf16), C.c_void_p(stream.cuda_stream)), "lg_process_batch")
            for k in range(1, len(parts)):
                cur.wait_stream(self._lane_streams[k - 1])
        return self._records(res, N.FRAME_RESULT) if sync else res

    def lane_split(self, n: int):
        """[(lo, hi)] frame ranges of the parts a batch of n frames is processed in (one range without lanes)."""
        lanes = 1 + len(self._lane_ctx) if self.lanes_active else 1
        if lanes == 1 or n < 2 * lanes:
            return [(0, n)]
        per = -(-n // lanes)
        return [(lo, min(n, lo + per)) for lo in range(0, n, per)]

    def _host_frames(self, t, dtype, name):
        t = torch.as_tensor(t)
        if t.dim() == 2:
            t = t.unsqueeze(0)
        if t.device.type != "cpu":
            raise ValueError(f"{name}: process_batch_host takes HOST tensors (got {t.device}); use process_batch")
        if t.dtype != dtype:
            raise ValueError(f"{name}: expected dtype {dtype}, got {t.dtype}")
        if t.dim() != 3 or tuple(t.shape[1:]) != (self.H, self.W):
            raise ValueError(f"{name}: expected [n, {self.H}, {self.W}], got {tuple(t.shape)}")
        if not t.is_contiguous():
            raise ValueError(f"{name}: must be contiguous")
        if t.shape[0] < 1 or t.shape[0] > self.B:
            raise ValueError(f"{name}: {t.shape[0]} frames, the engine was built for 1..{self.B}")
        return t

    def set_record_output(self, records: "torch.Tensor | None"):
        """Device float32 tensor [>= frames, 20, 4] that every following process_batch / process_batch_host call fills
        with the frames' candidate records (x, y, traditional score, ML score), written by the fusion kernel; None = off.
        With lanes, part k writes at its frame offset.  This is the send buffer of dist.gather_candidate_records."""
        if records is not None:
            if (not records.is_cuda or records.dtype != torch.float32 or not records.is_contiguous()
                    or records.dim() != 3 or tuple(records.shape[1:]) != (N.TOP_K, 4)):
                raise ValueError("records: expected a contiguous CUDA float32 tensor [frames, 20, 4]")
        self._rec_out = records
        if records is None or not self._lane_ctx:
            N.check(self.lib.lg_set_record_output(self._ctx, _ptr(records)), "lg_set_record_output")
            for ctx in self._lane_ctx:
                N.check(self.lib.lg_set_record_output(ctx, None), "lg_set_record_output")

    def process_batch_host(self, labels_host, depth_host, cam: N.Camera, use_bf16: bool = False):
        """Pinned (or plain) HOST tensors int16 / float32 [n, H, W] in, structured ndarray out; copies are inside the call."""
        labels_host = self._host_frames(labels_host, torch.int16, "labels")
        depth_host = self._host_frames(depth_host, torch.float32, "depth")
        n = labels_host.shape[0]
        if depth_host.shape[0] != n:
            raise ValueError("labels and depth hold different numbers of frames")
        if self._lane_ctx and self._rec_out is not None:
            N.check(self.lib.lg_set_record_output(self._ctx, _ptr(self._rec_out)), "lg_set_record_output")
        out = np.empty(n, dtype=N.FRAME_RESULT)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_process_batch_host(self._ctx, _ptr(labels_host), _ptr(depth_host), n, C.byref(cam),
                                                   out.ctypes.data_as(C.c_void_p), int(use_bf16), _stream()),
                    "lg_process_batch_host")
        return out

    def select_leaf(self, labels, depth, cam: N.Camera):
        labels = self._frames(labels, torch.int16, "labels")
        depth = self._frames(depth, torch.float32, "depth")
        n = labels.shape[0]
        ids = torch.empty(n, dtype=torch.int32, device=self.device)
        rec = torch.empty(n * self.L * N.LEAF_RECORD.itemsize, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_select_leaf(self._ctx, _ptr(labels), _ptr(depth), n, C.byref(cam), _ptr(ids), _ptr(rec),
                                            _stream()), "lg_select_leaf")
        return ids.cpu().numpy(), self._records(rec, N.LEAF_RECORD).reshape(n, self.L)

    def select_grasp_point(self, mask, depth, cam: N.Camera, use_bf16: bool = False):
        mask = self._frames(mask, torch.uint8, "mask")
        depth = self._frames(depth, torch.float32, "depth")
        n = mask.shape[0]
        res = self._new_results(n)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_select_grasp_point(self._ctx, _ptr(mask), _ptr(depth), n, C.byref(cam), _ptr(res),
                                                   int(use_bf16), _stream()), "lg_select_grasp_point")
        return self._records(res, N.FRAME_RESULT)

    # ---- training samples (data_collector.py:175-348) ---------------------------------------------------
    def collect_samples(self, depth, labels=None, mask=None, seed: int = 0, first_frame_index: int = 0, grasp_xy=None,
                        total_score=None):
        """Samples of the batch this engine has just processed: call right after process_batch (pass the same labels)
        or select_grasp_point (pass the same mask).  Returns (patches float32 device [n,7,9,32,32], meta structured
        ndarray [n,7] (SAMPLE_META), set_sizes int32 ndarray [n,3])."""
        if (labels is None) == (mask is None):
            raise ValueError("pass exactly one of labels / mask")
        if self._lane_ctx and self.lanes_active:
            raise N.NativeError("collect_samples needs the whole batch in one context: build the engine with lanes=1")
        depth = self._frames(depth, torch.float32, "depth")
        src = self._frames(labels, torch.int16, "labels") if labels is not None else self._frames(mask, torch.uint8, "mask")
        n = depth.shape[0]
        if grasp_xy is not None:
            grasp_xy = torch.as_tensor(np.asarray(grasp_xy, dtype=np.int32).reshape(n, 2)).to(self.device).contiguous()
        if total_score is not None:
            total_score = torch.as_tensor(np.asarray(total_score, dtype=np.float64).reshape(n)).to(self.device).contiguous()
        patches = torch.zeros(n, N.SAMPLES_PER_FRAME, N.CHANNELS, N.PATCH, N.PATCH, dtype=torch.float32, device=self.device)
        meta = torch.zeros(n * N.SAMPLES_PER_FRAME * N.SAMPLE_META.itemsize, dtype=torch.uint8, device=self.device)
        sizes = torch.zeros(n, 3, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_collect_samples(self._ctx, _ptr(src if labels is not None else None),
                                                _ptr(src if mask is not None else None), _ptr(depth), n,
                                                C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_uint64(int(first_frame_index)),
                                                _ptr(grasp_xy), _ptr(total_score), _ptr(patches), _ptr(meta), _ptr(sizes),
                                                _stream()), "lg_collect_samples")
        return patches, self._records(meta, N.SAMPLE_META).reshape(n, N.SAMPLES_PER_FRAME), sizes.cpu().numpy()

    def collector_points(self, kind: int, ranks, labels=None, mask=None) -> np.ndarray:
        """Points (x, y) of a candidate set by rank in the reference's list order, after collect_samples on the same
        batch: kind 0 tip, 1 stem, 2 edge; ranks [n, nq] are taken modulo the set size."""
        if (labels is None) == (mask is None):
            raise ValueError("pass exactly one of labels / mask")
        src = self._frames(labels, torch.int16, "labels") if labels is not None else self._frames(mask, torch.uint8, "mask")
        n = src.shape[0]
        ranks = torch.as_tensor(np.asarray(ranks, dtype=np.uint32).reshape(n, -1).astype(np.int64)).to(torch.int32)
        ranks = ranks.to(self.device).contiguous()
        nq = ranks.shape[1]
        xy = torch.empty(n, nq, 2, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_collector_points(self._ctx, _ptr(src if labels is not None else None),
                                                 _ptr(src if mask is not None else None), n, int(kind), _ptr(ranks), nq,
                                                 _ptr(xy), _stream()), "lg_collector_points")
        return xy.cpu().numpy()

    def last_patches(self, n: int) -> torch.Tensor:
        out = torch.empty(n, N.TOP_K, N.CHANNELS, N.PATCH, N.PATCH, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_patches(self._ctx, _ptr(out), n, _stream()), "lg_patches")
        return out

    # ---- stages ----------------------------------------------------------------------------------------
    def chamfer(self, mask, invert: bool = False, want_q16: bool = True):
        mask = self._frames(mask, torch.uint8, "mask")
        n = mask.shape[0]
        dist = torch.empty(n, self.H, self.W, dtype=torch.float32, device=self.device)
        q16 = torch.empty(n, self.H, self.W, dtype=torch.int32, device=self.device) if want_q16 else None
        mx = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_chamfer_transform(self._ctx, _ptr(mask), n, int(invert), _ptr(dist), _ptr(q16), _ptr(mx),
                                                  _stream()), "lg_chamfer_transform")
        return dist, q16, mx

    def edt_squared(self, mask, argmax_only: bool = False):
        """Exact squared EDT and its first arg-max; argmax_only=True takes the pruned search the leaf-selection
        stage uses (no field is written) and returns (None, argmax)."""
        mask = self._frames(mask, torch.uint8, "mask")
        n = mask.shape[0]
        d2 = None if argmax_only else torch.empty(n, self.H, self.W, dtype=torch.int32, device=self.device)
        am = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_edt_squared(self._ctx, _ptr(mask), n, _ptr(d2), _ptr(am), _stream()), "lg_edt_squared")
        return d2, am

    def score_maps(self, mask, depth, cam: N.Camera):
        """Full-frame maps with the reference's names and dtypes (device tensors) + 'valid' + 'angle'."""
        mask = self._frames(mask, torch.uint8, "mask")
        depth = self._frames(depth, torch.float32, "depth")
        n = mask.shape[0]
        f64 = lambda: torch.empty(n, self.H, self.W, dtype=torch.float64, device=self.device)
        f32 = lambda: torch.empty(n, self.H, self.W, dtype=torch.float32, device=self.device)
        out = {"sdf_score": f64(), "approach_score": f64(), "flatness_map": f32(), "isolation_map": f64(),
               "distance_map": f32(), "accessibility_map": f64(), "stem_penalty": f32(), "traditional_score": f64()}
        valid = torch.empty(n, self.H, self.W, dtype=torch.uint8, device=self.device)
        angle = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_score_maps(
                self._ctx, _ptr(mask), _ptr(depth), n, C.byref(cam), _ptr(out["sdf_score"]), _ptr(out["approach_score"]),
                _ptr(out["flatness_map"]), _ptr(out["isolation_map"]), _ptr(out["distance_map"]),
                _ptr(out["accessibility_map"]), _ptr(out["stem_penalty"]), _ptr(out["traditional_score"]), _ptr(valid),
                _ptr(angle), _stream()), "lg_score_maps")
        out["valid"] = valid
        out["angle"] = angle
        return out

    def candidate_points(self, score, valid):
        score = self._frames(score, torch.float64, "score")
        valid = self._frames(valid, torch.uint8, "valid")
        n = score.shape[0]
        xy = torch.empty(n, N.TOP_K, 2, dtype=torch.int32, device=self.device)
        cnt = torch.empty(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_candidate_points(self._ctx, _ptr(score), _ptr(valid), n, _ptr(xy), _ptr(cnt), _stream()),
                    "lg_candidate_points")
        return xy, cnt

    def leaf_orientation(self, mask):
        mask = self._frames(mask, torch.uint8, "mask")
        n = mask.shape[0]
        out = torch.empty(n, 5, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self.lib.lg_leaf_orientation(self._ctx, _ptr(mask), n, _ptr(out), _stream()), "lg_leaf_orientation")
        return out
