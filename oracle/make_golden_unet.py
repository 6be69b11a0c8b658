"""Pin oracle/unetpp_ref.py against the reference's OWN module and freeze a small fixture (build container only).

    python -m oracle.make_golden_unet      # writes tests/golden/unetpp.json

Imports /root/reference/src/preprocessing/segmentation/model.py unmodified, checks that the oracle has the same
state_dict keys, draws the same initial parameters from the same seed, and produces bit-identical logits on the CPU with
the oracle's (randomised-BatchNorm) parameters loaded into the reference module; then stores sample logits."""
from __future__ import annotations

import hashlib
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def main():
    sys.path.insert(0, REPO)
    sys.path.insert(0, "/root/reference")
    from oracle.unetpp_ref import NestedUNetRef, seeded_input, seeded_model
    from src.preprocessing.segmentation.model import FingerprintSegmentationModel, NestedUNet
    torch.manual_seed(0); a = NestedUNetRef()
    torch.manual_seed(0); b = NestedUNet()
    assert list(a.state_dict()) == list(b.state_dict())
    assert all(torch.equal(a.state_dict()[k], b.state_dict()[k]) for k in a.state_dict())
    ora = seeded_model(0)
    ref = FingerprintSegmentationModel().eval()
    ref.model.load_state_dict(ora.state_dict(), strict=True)
    cases = []
    for seed, n, h, w in ((0, 1, 32, 48), (1, 2, 64, 64)):
        x = seeded_input(seed, n, h, w)
        with torch.no_grad():
            yo, yr = ora(x), ref(x)
        assert torch.equal(yo, yr), "oracle != reference module"
        flat = yr.flatten()
        idx = torch.linspace(0, flat.numel() - 1, 16).long()
        cases.append({"seed": seed, "n": n, "h": h, "w": w, "sum": float(flat.double().sum()), "absmax": float(flat.abs().max()),
                      "sample_index": idx.tolist(), "sample": [float(v) for v in flat[idx]],
                      "sha256_float32": hashlib.sha256(yr.numpy().tobytes()).hexdigest()})
    out = os.path.join(REPO, "tests", "golden", "unetpp.json")
    with open(out, "w") as f:
        json.dump({"keys": len(a.state_dict()), "cases": cases}, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
