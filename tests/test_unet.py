"""U-Net++ segmenter inference: the torch fp32 oracle against the fixture made from the reference's own
module (CPU), and the tcgen05 engine against the oracle (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.unetpp_ref import seeded_input, seeded_model

TOL = 2e-4          # |logit - reference| <= TOL * max|reference| : fp32 round-off of ~25 chained convolutions (3xTF32 GEMMs)


def test_oracle_reproduces_reference_fixture():
    g = json.load(open(os.path.join(GOLDEN, "unetpp.json")))
    m = seeded_model(0)
    assert len(m.state_dict()) == g["keys"]
    for c in g["cases"]:
        with torch.no_grad():
            y = m(seeded_input(c["seed"], c["n"], c["h"], c["w"])).flatten()
        got = y[torch.tensor(c["sample_index"])].numpy()
        assert np.allclose(got, c["sample"], rtol=0, atol=1e-5 * c["absmax"]), (got, c["sample"])     # other CPUs: other conv kernels
        assert abs(float(y.double().sum()) - c["sum"]) <= 1e-5 * c["absmax"] * y.numel()


def test_mirror_has_the_reference_surface():
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.segmentation import inference, model
    assert hasattr(model, "NestedUNet") and hasattr(model, "FingerprintSegmentationModel")
    assert hasattr(inference, "preprocess_image") and hasattr(inference, "mask_to_rgb")
    logits = np.array([[-2.0, 3.0], [0.1, -0.1]], np.float32)
    rgb = np.full((4, 4, 3), 100, np.uint8)
    mask, seg, ov = inference.mask_to_rgb(logits, rgb)                    # inference.py:95-110 semantics
    assert mask.shape == (4, 4) and set(np.unique(mask)) == {0, 255} and mask[0, 0] == 0 and mask[0, 3] == 255
    assert (seg[mask == 0] == 0).all() and (seg[mask == 255] == 100).all() and ov.shape == rgb.shape


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w", [(1, 32, 48), (2, 64, 64), (2, 224, 224)])
def test_engine_matches_torch_fp32(n, h, w):
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.segmentation.model import FingerprintSegmentationModel
    ora = seeded_model(0)
    x = seeded_input(7, n, h, w)
    with torch.no_grad():
        want = ora(x).numpy()
    m = FingerprintSegmentationModel(image_size=(h, w), max_batch=2)
    m.load_state_dict({"model." + k: v for k, v in ora.state_dict().items()})     # keys as the reference's wrapper saves them
    m.eval()
    got = m(x.numpy())
    assert got.shape == want.shape == (n, 1, h, w)
    scale = float(np.abs(want).max())
    err = float(np.abs(got - want).max())
    assert err <= TOL * scale, f"max |diff| {err:.3e} vs scale {scale:.3e}"
    sure = np.abs(want) > 10 * TOL * scale                               # away from the decision boundary the masks are identical
    assert np.array_equal((got > 0)[sure], (want > 0)[sure])
    total, tc = m.model.launches()
    assert tc == 20 and total > tc                                       # ten evaluated ConvBlocks x two tensor-core convolutions
    again = m(torch.from_numpy(x.numpy()))                               # torch in -> torch out, deterministic
    assert isinstance(again, torch.Tensor) and np.array_equal(again.numpy(), got)


@pytest.mark.gpu
def test_engine_requires_parameters_and_valid_shapes():
    from multimodal_biometric_fingerprints_palms_b200 import FpbError
    from multimodal_biometric_fingerprints_palms_b200.preprocessing.segmentation.model import NestedUNet
    m = NestedUNet()
    with pytest.raises(RuntimeError, match="load_state_dict"):
        m(np.zeros((1, 3, 32, 32), np.float32))
    m.load_state_dict(seeded_model(0).state_dict())
    with pytest.raises(FpbError, match="multiples of 16"):
        m(np.zeros((1, 3, 30, 32), np.float32))
    with pytest.raises(RuntimeError, match="Missing key"):
        NestedUNet().load_state_dict({"final.bias": np.zeros(1, np.float32)})
