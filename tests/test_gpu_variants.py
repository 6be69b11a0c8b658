"""Alternative kernels behind the same stage entry points, each forced through its environment switch in a child process
(the switches are read once per process): every variant must reproduce the oracle exactly as the default one does.

  * NLM (K2): k_nlm_sym with several vertical segments per image, without TMA; the older kernels k_nlm3 / k_nlm.
  * K4 tail / K7 component filters: k_bin_finish_cl on clusters of 2 and 8 CTAs at every image size (by default it only takes
    images that are too large for one CTA), with the band-local and with the flat (all-global) union-find.

Run on the B200 box:  python -m pytest tests -m gpu -q"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

NLM_CHILD = r"""
import sys, json
sys.path.insert(0, %r)
import cv2, numpy as np
from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline, synth
rng = np.random.default_rng(3)
bad = {}
for (h, w, n, kind) in [(320, 240, 3, "ridge"), (131, 97, 2, "ridge"), (64, 48, 2, "noise"), (96, 480, 1, "noise"),
                        (300, 5, 1, "noise"), (5, 300, 1, "noise"), (37, 241, 1, "noise"), (512, 512, 1, "ridge")]:
    imgs = (np.stack([synth.ridge_image(h, w, seed=40 + i, period=None) for i in range(n)]) if kind == "ridge"
            else rng.integers(0, 256, (n, h, w), dtype=np.uint8))
    p = FingerprintPipeline(h, w, max_batch=n)
    _, nlm = p.denoise(imgs, with_nlm=True)
    p.close()
    bad["%%dx%%d" %% (h, w)] = sum(int((nlm[i] != cv2.fastNlMeansDenoising(imgs[i], None, 10, 7, 21)).sum()) for i in range(n))
print(json.dumps(bad))
""" % ROOT


def _child_env(extra):
    env = dict(os.environ)
    for k in ("FPB_NLM_V", "FPB_NLM_MMA", "FPB_NLM_SEGS", "FPB_NO_TMA", "FPB_NLM_THREADS", "FPB_BIN_CLUSTER", "FPB_CLUSTER_FLAT",
              "FPB_NO_CLUSTER"):
        env.pop(k, None)
    env.update(extra)
    return env


@pytest.mark.parametrize("env", [{}, {"FPB_NLM_SEGS": "3"}, {"FPB_NLM_SEGS": "16"}, {"FPB_NO_TMA": "1"}, {"FPB_NLM_THREADS": "256"},
                                 {"FPB_NLM_V": "3"}, {"FPB_NLM_V": "1"}],
                         ids=["sym", "sym_3_segments", "sym_16_segments", "sym_no_tma", "sym_256_threads", "k_nlm3", "k_nlm"])
def test_nlm_kernel_variants_bit_exact_against_opencv(env):
    import json
    r = subprocess.run([sys.executable, "-c", NLM_CHILD], env=_child_env(env), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    bad = json.loads(r.stdout.strip().splitlines()[-1])
    assert all(v == 0 for v in bad.values()), bad


@pytest.mark.parametrize("env", [{"FPB_BIN_CLUSTER": "8"}, {"FPB_BIN_CLUSTER": "2"}, {"FPB_BIN_CLUSTER": "4", "FPB_CLUSTER_FLAT": "1"},
                                 {"FPB_NO_CLUSTER": "1"}],
                         ids=["cluster8", "cluster2", "cluster4_flat_union_find", "no_cluster"])
def test_cluster_kernels_reproduce_the_component_stages(env):
    """The K4 / K7 / end-to-end parity tests of the suite, re-run with the cluster kernel forced (or forbidden)."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_stages.py"),
                        os.path.join(ROOT, "tests", "test_gpu_pipeline.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "k4 or k7 or config2 or config4 or highres or fused_batch or odd_shapes or degenerate"],
                       env=_child_env(env), capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-2000:]
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in tail and "failed" not in tail, tail
