"""GraspPointCNN drop-in (reference scripts/utils/ml_grasp_optimizer/model.py:5-128).

The module keeps the reference's constructor, sub-module layout and parameter names, so ``load_state_dict`` of a
reference checkpoint (``checkpoint['model_state_dict']``, grasp_point_selector.py:48-49) works unchanged for every
architecture the reference can build (attention 'spatial' / 'channel' / 'hybrid' / 'none', any encoder filter list of
the sweep in train_model_mlflow.py:173-182).  ``forward`` in eval mode runs the hand-written CUDA kernels
(csrc/lg_cnn*.cu) on BatchNorm-folded weights; there is no torch / cuDNN forward behind it.  The default architecture
(the one the live node builds) and its attention variants (same encoder [64, 128, 256]) can run on the tensor cores
(``use_bf16``); the other encoders use the fp32 kernels.
Training is out of scope: calling it in training mode raises.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _native as N

_BN_EPS = 1e-5


def _block(cin, cout):
    # indices 0,1 / 3,4 hold conv+bn exactly like the reference's Sequential (state_dict keys)
    return nn.Sequential(
        nn.Conv2d(cin, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        nn.Conv2d(cout, cout, 3, padding=1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
        nn.MaxPool2d(2), nn.Dropout2d(0.3))


def _channel_attention(c):
    return nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(c, c // 16, 1), nn.ReLU(inplace=True), nn.Conv2d(c // 16, c, 1),
                         nn.Sigmoid())


def architecture_of(sd: dict):
    """(attention_type, encoder_filters) recovered from a state_dict's keys and shapes."""
    n_blocks = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("encoder."))
    filters = [int(sd[f"encoder.{b}.0.weight"].shape[0]) for b in range(n_blocks)]
    if "spatial_attention.0.weight" in sd:
        att = "hybrid"
    elif "attention.0.weight" in sd:
        att = "spatial"
    elif "attention.1.weight" in sd:
        att = "channel"
    else:
        att = "none"
    return att, filters


def fold_batchnorm(sd: dict) -> dict:
    """Fold eval-mode BatchNorm into the preceding conv / linear.  Returns {'convs': [(w [Cout,Cin,3,3], b)] (two per
    block), 'spatial': (w [1,C,1,1], b) | None, 'channel': ((w1 [C/16,C,1,1], b1), (w2 [C,C/16,1,1], b2)) | None,
    'fcs': [(w [out,in], b)] x 4}, all float64."""
    g = lambda k: sd[k].detach().double().cpu().numpy()

    def fold(wk, bk, bn):
        w, b = g(wk), g(bk)
        if bn is not None:
            s = g(bn + ".weight") / np.sqrt(g(bn + ".running_var") + _BN_EPS)
            w = w * s.reshape((-1,) + (1,) * (w.ndim - 1))
            b = (b - g(bn + ".running_mean")) * s + g(bn + ".bias")
        return w, b

    att, filters = architecture_of(sd)
    out = {"convs": [], "spatial": None, "channel": None, "fcs": [], "attention_type": att, "encoder_filters": filters}
    for blk in range(len(filters)):
        out["convs"].append(fold(f"encoder.{blk}.0.weight", f"encoder.{blk}.0.bias", f"encoder.{blk}.1"))
        out["convs"].append(fold(f"encoder.{blk}.3.weight", f"encoder.{blk}.3.bias", f"encoder.{blk}.4"))
    if att == "spatial":
        out["spatial"] = fold("attention.0.weight", "attention.0.bias", None)
    elif att == "channel":
        out["channel"] = (fold("attention.1.weight", "attention.1.bias", None), fold("attention.3.weight", "attention.3.bias", None))
    elif att == "hybrid":
        out["spatial"] = fold("spatial_attention.0.weight", "spatial_attention.0.bias", None)
        out["channel"] = (fold("channel_attention.1.weight", "channel_attention.1.bias", None),
                          fold("channel_attention.3.weight", "channel_attention.3.bias", None))
    for lin, bn in ((0, "classifier.1"), (4, "classifier.5"), (8, "classifier.9"), (12, None)):
        out["fcs"].append(fold(f"classifier.{lin}.weight", f"classifier.{lin}.bias", bn))
    return out


def pack_weights(sd: dict) -> np.ndarray:
    """Folded weights as the float32 blob csrc/lg_cnn.cu reads:
    conv l: w[ky][kx][Cin][Cout], b[Cout];  spatial attention w[C], b[1];  channel attention w1[C][C/16], b1, w2[C/16][C], b2;
    fc k: w[in][out], b[out]."""
    f = fold_batchnorm(sd)
    parts = []
    for w, b in f["convs"]:
        parts += [np.transpose(w, (2, 3, 1, 0)).ravel(), b.ravel()]
    if f["spatial"] is not None:
        parts += [f["spatial"][0].ravel(), f["spatial"][1].ravel()]
    if f["channel"] is not None:
        (w1, b1), (w2, b2) = f["channel"]
        parts += [np.transpose(w1[:, :, 0, 0], (1, 0)).ravel(), b1.ravel(), np.transpose(w2[:, :, 0, 0], (1, 0)).ravel(), b2.ravel()]
    for w, b in f["fcs"]:
        parts += [np.transpose(w, (1, 0)).ravel(), b.ravel()]
    return np.ascontiguousarray(np.concatenate(parts).astype(np.float32))


class GraspPointCNN(nn.Module):
    """Same constructor and sub-modules as the reference (model.py:6-86)."""

    def __init__(self, in_channels=9, attention_type="spatial", encoder_filters=(64, 128, 256)):
        super().__init__()
        if in_channels != 9:
            raise NotImplementedError("the CUDA path consumes the 9-channel patch tensor of grasp_point_selector.py:127")
        self.attention_type = attention_type
        self.encoder_filters = list(encoder_filters)
        self._config = N.cnn_config(attention_type if attention_type in N.ATTENTION_CODES else "none", self.encoder_filters)
        self.encoder = nn.ModuleList()
        cin = in_channels
        for f in self.encoder_filters:
            self.encoder.append(_block(cin, f))
            cin = f
        if attention_type == "spatial":
            self.attention = nn.Sequential(nn.Conv2d(cin, 1, 1), nn.Sigmoid())
        elif attention_type == "channel":
            self.attention = _channel_attention(cin)
        elif attention_type == "hybrid":
            self.spatial_attention = nn.Sequential(nn.Conv2d(cin, 1, 1), nn.Sigmoid())
            self.channel_attention = _channel_attention(cin)
        else:   # 'none' (the reference treats every other string like this, model.py:58-59)
            self.attention = None
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

    @property
    def is_default_architecture(self) -> bool:
        return self.attention_type == "spatial" and self.encoder_filters == [64, 128, 256]

    @property
    def has_tensor_core_path(self) -> bool:
        """The tcgen05 convolutions are built for the encoder [64, 128, 256]; the attention type only changes the fp32 tail."""
        return self.encoder_filters == [64, 128, 256]

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
            self._engine.set_cnn_weights(self.packed(), None if self.is_default_architecture else self._config)
            self._packed_version = ver
        xin = x.detach().to(dev, torch.float32).contiguous()
        return self._engine.cnn_forward(xin, self.use_bf16 and self.has_tensor_core_path).reshape(-1, 1).to(x.device)
