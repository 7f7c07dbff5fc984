"""ImageProcessor drop-in (reference scripts/utils/image_processor.py:8-64).

On the B200 path the Gaussian / Sobel stencils live inside the fused score kernel (csrc/lg_score.cu), so
``GraspPointSelector`` here does not call this class; it carries the constructor arguments, the kernel
tensors the reference exposes through ``get_kernel``, and ``smooth_depth`` (image_processor.py:56-64) as a
stand-alone device kernel (``lg_smooth_depth``) for callers that use it on its own.
"""
from __future__ import annotations

import colorsys

import ctypes as C

import numpy as np
import torch

from . import _native as N


class ImageProcessor:
    def __init__(self, height, width, kernel_size, gaussian_size):
        if gaussian_size != 5:
            raise NotImplementedError("the fused flatness kernel implements the reference's 5x5 Gaussian "
                                      "(leaf_grasp_node_v3.py:37)")
        self.height, self.width = height, width
        sob = torch.tensor([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], dtype=torch.float32)
        self.kernels = {"isolation": torch.ones(kernel_size, kernel_size), "sobel_x": sob, "sobel_y": sob.t(),
                        "gaussian": self._gaussian(gaussian_size)}
        self.color_map = {}

    @staticmethod
    def _gaussian(size):
        sigma, c = size / 6.0, size // 2
        x, y = np.meshgrid(np.arange(size), np.arange(size))
        k = np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2))
        return torch.tensor(k / k.sum(), dtype=torch.float32)

    def get_kernel(self, name, device):
        k = self.kernels.get(name)
        return None if k is None else k.to(device)

    def generate_color(self, leaf_id):
        if leaf_id not in self.color_map:
            rgb = colorsys.hsv_to_rgb((leaf_id * 0.618033988749895) % 1.0, 0.8, 0.95)
            self.color_map[leaf_id] = tuple(int(255 * v) for v in rgb)
        return self.color_map[leaf_id]

    def calculate_centroid(self, leaf_mask):
        ys, xs = torch.where(leaf_mask)
        return float(xs.float().mean()), float(ys.float().mean())

    def smooth_depth(self, depth_patch, device):
        """Reflect-pad by 2 + 5x5 Gaussian of a 2-D float image (image_processor.py:56-64) on the CUDA device; returns a
        float32 tensor [h, w] on that device.  No CPU fallback: raises without a GPU."""
        if not torch.cuda.is_available():
            raise N.NativeError("ImageProcessor.smooth_depth needs a CUDA device; there is no CPU fallback")
        dev = torch.device(device)
        if dev.type != "cuda":
            dev = torch.device("cuda", torch.cuda.current_device())
        x = torch.as_tensor(depth_patch).to(dev, torch.float32).contiguous()
        if x.dim() != 2:
            raise ValueError(f"smooth_depth expects a 2-D image, got {tuple(x.shape)}")
        h, w = x.shape
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            N.check(N.lib().lg_smooth_depth(C.c_void_p(x.data_ptr()), 1, h, w, C.c_void_p(out.data_ptr()),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)), "lg_smooth_depth")
        return out
