"""Golden vectors for the GraspPointCNN architecture variants, produced by the
REFERENCE's own model class (scripts/utils/ml_grasp_optimizer/model.py) in the build container:

    python tests/golden/make_cnn_variants.py

For every variant of the sweep (train_model_mlflow.py:173-182: attention spatial / channel / hybrid / none, filter
lists [32,64,128], [64,128,256], [64,128,256,512], [128,256,512]) it stores the state_dict key -> shape table, 6 seeded
patches and the reference's logits for weights rebuilt from that table by oracle.seeded_state_dict_from_shapes.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference/scripts/utils/ml_grasp_optimizer")
import leafgrasp_oracle as O  # noqa: E402
from model import GraspPointCNN  # noqa: E402  (the reference's class)

VARIANTS = [("spatial", [32, 64, 128]), ("channel", [64, 128, 256]), ("hybrid", [64, 128, 256]), ("none", [64, 128, 256]),
            ("spatial", [64, 128, 256, 512]), ("channel", [128, 256, 512]), ("hybrid", [32, 64, 128]), ("none", [64, 128, 256, 512])]


def main():
    torch.set_num_threads(1)
    rng = np.random.default_rng(77)
    x = torch.from_numpy(rng.random((6, 9, 32, 32), dtype=np.float32))
    x[:, 1] = (x[:, 1] > 0.5).float()
    out = {"x_seed": 77, "variants": []}
    logits = {}
    for i, (att, filters) in enumerate(VARIANTS):
        net = GraspPointCNN(in_channels=9, attention_type=att, encoder_filters=filters)
        shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
        sd = O.seeded_state_dict_from_shapes(shapes, 500 + i)
        net.load_state_dict(sd)
        net.eval()
        with torch.no_grad():
            y = net(x).reshape(-1).numpy()
        out["variants"].append({"attention_type": att, "encoder_filters": filters, "seed": 500 + i, "shapes": shapes})
        logits[f"logits_{i}"] = y
        print(att, filters, y)
    json.dump(out, open(os.path.join(HERE, "cnn_variants.json"), "w"))
    np.savez_compressed(os.path.join(HERE, "cnn_variants.npz"), x=x.numpy(), **logits)


if __name__ == "__main__":
    main()
