"""GraspPointCNN drop-in (reference scripts/utils/ml_grasp_optimizer/model.py:5-128).

The module keeps the reference's parameter names, so ``load_state_dict`` of a reference checkpoint
(``checkpoint['model_state_dict']``, grasp_point_selector.py:48-49) works unchanged.  ``forward`` in eval
mode runs the hand-written CUDA kernels (csrc/lg_cnn*.cu) on BatchNorm-folded weights; there is no
torch / cuDNN forward behind it.  Training is out of scope: calling it in
training mode raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _native as N

_FILTERS = (64, 128, 256)
_BN_EPS = 1e-5


def _block(cin, cout):
    # indices 0,1 / 3,4 hold conv+bn exactly like the reference's Sequential (state_dict keys)
    return nn.Sequential(
        nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        nn.MaxPool2d(2), nn.Dropout2d(0.3))


def fold_batchnorm(sd: dict) -> list:
    """Fold eval-mode BatchNorm into the preceding conv / linear.  Returns, in network order, a list of
    (weight float64 ndarray, bias float64 ndarray): 6 convs [Cout,Cin,3,3], attention [1,256,1,1],
    4 linears [out,in]."""
    g = lambda k: sd[k].detach().double().cpu().numpy()
    out = []

    def fold(wk, bk, bn):
        w, b = g(wk), g(bk)
        if bn is not None:
            s = g(bn + ".weight") / np.sqrt(g(bn + ".running_var") + _BN_EPS)
            w = w * s.reshape((-1,) + (1,) * (w.ndim - 1))
            b = (b - g(bn + ".running_mean")) * s + g(bn + ".bias")
        out.append((w, b))

    for blk in range(3):
        fold(f"encoder.{blk}.0.weight", f"encoder.{blk}.0.bias", f"encoder.{blk}.1")
        fold(f"encoder.{blk}.3.weight", f"encoder.{blk}.3.bias", f"encoder.{blk}.4")
    fold("attention.0.weight", "attention.0.bias", None)
    for lin, bn in ((0, "classifier.1"), (4, "classifier.5"), (8, "classifier.9"), (12, None)):
        fold(f"classifier.{lin}.weight", f"classifier.{lin}.bias", bn)
    return out


def pack_weights(sd: dict) -> np.ndarray:
    """Folded weights as the float32 blob csrc/lg_cnn.cu reads:
    conv l: w[ky][kx][Cin][Cout], b[Cout];  attention w[256], b[1];  fc k: w[in][out], b[out]."""
    folded = fold_batchnorm(sd)
    parts = []
    for w, b in folded[:6]:
        parts += [np.transpose(w, (2, 3, 1, 0)).ravel(), b.ravel()]
    aw, ab = folded[6]
    parts += [aw.ravel(), ab.ravel()]
    for w, b in folded[7:]:
        parts += [np.transpose(w, (1, 0)).ravel(), b.ravel()]
    return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))


class GraspPointCNN(nn.Module):
    """Same constructor as the reference.  The CUDA path implements the configuration the reference's
    live code builds (``GraspPointCNN(in_channels=9)``: spatial attention, filters [64,128,256]); the
    sweep variants of model.py:30-60 are a SURVEY 8f "next" row and raise NotImplementedError."""

    def __init__(self, in_channels=9, attention_type="spatial", encoder_filters=(64, 128, 256)):
        super().__init__()
        if in_channels != 9 or attention_type != "spatial" or tuple(encoder_filters) != _FILTERS:
            raise NotImplementedError(
                "CUDA path covers GraspPointCNN(in_channels=9, attention_type='spatial', "
                "encoder_filters=[64,128,256]) only")
        self.attention_type = attention_type
        self.encoder_filters = list(encoder_filters)
        self.encoder = nn.ModuleList()
        cin = in_channels
        for f in encoder_filters:
            self.encoder.append(_block(cin, f))
            cin = f
        self.attention = nn.Sequential(nn.Conv2d(cin, 1, 1), nn.Sigmoid())
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential(
            nn.Linear(cin, cin), nn.BatchNorm1d(cin), nn.ReLU(inplace=True), nn.Dropout(0.5),
            nn.Linear(cin, cin // 2), nn.BatchNorm1d(cin // 2), nn.ReLU(inplace=True), nn.Dropout(0.5),
            nn.Linear(cin // 2, cin // 4), nn.BatchNorm1d(cin // 4), nn.ReLU(inplace=True), nn.Dropout(0.4),
            nn.Linear(cin // 4, 1))
        for m in self.modules():        # model.py:89-100
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight)
                nn.init.constant_(m.bias, 0)
        self._engine = None
        self._packed_version = None
        self.use_bf16 = False

    def _version(self):
        return tuple(int(t._version) for t in self.state_dict().values())

    def packed(self) -> np.ndarray:
        return pack_weights(self.state_dict())

    def forward(self, x):
        if self.training:
            raise RuntimeError("GraspPointCNN (B200 path) is inference only: call .eval() first")
        if not torch.cuda.is_available():
            raise N.NativeError("GraspPointCNN.forward needs a CUDA device; there is no CPU fallback")
        from .pipeline import GraspEngine
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if self._engine is None:
            self._engine = GraspEngine(1, 64, 64, 2, device=dev)
        ver = self._version()
        if ver != self._packed_version:
            self._engine.set_cnn_weights(self.packed())
            self._packed_version = ver
        xin = x.detach().to(dev, torch.float32).contiguous()
        return self._engine.cnn_forward(xin, self.use_bf16).reshape(-1, 1).to(x.device)
