"""CPU oracle for the fingerprint enhance -> minutiae hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`multimodal_biometric_fingerprints_palms_b200/`) imports this directory; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and there only as the checker / the timed CPU arm.

Layers (see DESIGN.md section "Oracle"):

* `oracle.skimage_compat`  - restatement of the five scikit-image functions the
  reference imports (`fingerprint_preprocess.py:5-6`); scikit-image is pinned
  `>=0.22,<1.0` by the reference's `config/environment.yml` but is absent from
  this image, so its published algorithms are restated on numpy / scipy.ndimage.
* `oracle.ref_pipeline`    - the reference's per-image pipeline restated on the
  same OpenCV / SciPy / NumPy calls the reference makes (those libraries ARE in
  this image, here and on the GPU box).  This is the parity oracle proper.
* `oracle.stages`          - numpy restatements of the *insides* of the OpenCV /
  SciPy calls (CLAHE, NLM, fixed-point Gaussian, box filter, gaussian_filter ...)
  - the specification the CUDA kernels were written from; every one is checked
  bit-for-bit (integer) or to float tolerance against the library in
  `tests/test_oracle_stages.py`.
* `oracle.ref_matching`    - the reference's RANSAC matcher (`src/matching/match.py`) restated on the same NumPy /
  scikit-learn calls, hypotheses consumed in seed order; pinned by `oracle/make_golden_matching.py`.
* `oracle.jpeg_idct`       - libjpeg's "islow" inverse DCT restated in NumPy; pinned against `cv2.imdecode`.
* `oracle.gabor_ext`       - NumPy statement of the EXTENSION rows G1/G2 (not in the reference: parity unpinned).

The sequential geometry routines of the CUDA library (contour tracing, convex hull, polygon fill ...) are checked on the
CPU by `tests/hostcheck` (a g++ build of the `FPB_HD` device routines against OpenCV), not by a separate C oracle.

Pinning: the reference ships no tests and no golden vectors, so
the oracle is pinned against outputs of the *reference's own modules* imported
from `/root/reference` in the build container (`oracle/make_golden.py`, fixtures
under `tests/golden/`).  The five scikit-image functions could not be run here:
for those, and in particular for the 256-entry skeletonize table, parity with
scikit-image itself is UNPINNED (see `oracle/skimage_compat.py` header).
"""
