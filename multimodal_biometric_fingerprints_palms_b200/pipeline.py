"""Batched host-side driver of the CUDA hot path (one handle = one GPU + one stream)."""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional

import numpy as np

from . import _native as N

TYPE_NAMES = ("ending", "bifurcation")
# fpb_minutia (include/fpb200.h) as a NumPy record: what `result_block` returns for a whole batch
MINUTIA_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("type", "<i4"), ("_pad", "<i4"), ("orientation", "<f8"),
                          ("quality", "<f8"), ("coherence", "<f8"), ("angular_stability", "<f8")])
assert MINUTIA_DTYPE.itemsize == C.sizeof(N.Minutia)


def _u8(a) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint8:
        raise TypeError(f"expected uint8 image data, got {a.dtype}")
    return a


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class FingerprintPipeline:
    """`max_batch` images of `height` x `width` uint8 pixels per call, on CUDA device `device`.

    `stream`: optional raw cudaStream_t (int), e.g. `torch.cuda.current_stream().cuda_stream`;
    by default the handle owns a non-blocking stream.  Not thread-safe: use one instance per thread.
    """

    def __init__(self, height: int, width: int, max_batch: int = 1, device: int = 0, stream: Optional[int] = None,
                 handoff: str = "file"):
        self._lib = N.load()
        self._h = C.c_void_p()
        self.H, self.W, self.max_batch, self.device = int(height), int(width), int(max_batch), int(device)
        rc = self._lib.fpb_create(C.byref(self._h), self.device, self.max_batch, self.H, self.W,
                                  C.c_void_p(stream) if stream else None)
        if rc != 0:
            raise N.FpbError(f"fpb_create failed ({rc}): {self._lib.fpb_last_error(None).decode()}")
        self.last_n = 0
        self.raw_capacity = int(self._lib.fpb_raw_capacity(self._h))
        self.set_handoff(handoff)
        self.rel_threshold = 0.1
        # scikit-image parity: an explicit table file wins; then, where scikit-image is installed
        # (the reference's own environment), a one-time self-check of the five restated functions against it.
        table = thin_table_from_env()
        if table is not None:
            self.set_thin_table(table)
        from . import selfcheck
        selfcheck.run_once(self)

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.fpb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc < 0:
            raise N.FpbError(f"{what} failed ({rc}): {self._lib.fpb_last_error(self._h).decode()}")
        return rc

    def _batch(self, a, dtype=np.uint8) -> np.ndarray:
        a = np.ascontiguousarray(a)
        if a.dtype != dtype:
            raise TypeError(f"expected {np.dtype(dtype)}, got {a.dtype}")
        if a.ndim == 2:
            a = a[None]
        if a.ndim != 3 or a.shape[1:] != (self.H, self.W):
            raise ValueError(f"expected [n,{self.H},{self.W}], got {a.shape}")
        if not 1 <= a.shape[0] <= self.max_batch:
            raise ValueError(f"batch {a.shape[0]} outside [1,{self.max_batch}]")
        return a

    @property
    def launch_count(self) -> int:
        return int(self._lib.fpb_launch_count(self._h))

    STAGE_NAMES = ("K1_normalize", "K2_denoise", "K3_segment", "K4_binarize", "K5_orientation", "K6_smooth",
                   "K7K8_thin_extract", "K9_postprocess", "nlm_kernel")

    def set_profiling(self, on: bool = True):
        self._ck(self._lib.fpb_set_profiling(self._h, int(bool(on))), "fpb_set_profiling")

    def stage_times_ms(self) -> Dict[str, float]:
        ms = (C.c_float * 16)()
        n = self._ck(self._lib.fpb_stage_times(self._h, ms, 16), "fpb_stage_times")
        return {self.STAGE_NAMES[i]: float(ms[i]) for i in range(n)}

    def kernel_times(self, enable: bool = True):
        """[(site, ms)] per kernel launch of the last run (site = "<file>:<line>" of the launch); also (re)arms recording."""
        buf = C.create_string_buffer(1 << 16)
        n = self._ck(self._lib.fpb_kernel_times(self._h, int(enable), buf, len(buf)), "fpb_kernel_times")
        out = []
        for ln in buf.raw[:n].decode().splitlines():
            site, ms = ln.rsplit(" ", 1)
            out.append((site, float(ms)))
        return out

    def sync(self):
        self._ck(self._lib.fpb_sync(self._h), "fpb_sync")

    def set_thin_table(self, table):
        t = _u8(np.asarray(table, dtype=np.uint8).reshape(256))
        if t.max(initial=0) > 3:
            raise ValueError("thinning table entries must be 0..3")
        self._ck(self._lib.fpb_set_thin_table(self._h, _ptr(t)), "fpb_set_thin_table")

    def set_rel_threshold(self, rel_thresh: float = 0.1):
        """`rel_thresh` of thinning_and_cleaning (fingerprint_preprocess.py:161): 0.1 on the reference's path."""
        self._ck(self._lib.fpb_set_rel_threshold(self._h, float(rel_thresh)), "fpb_set_rel_threshold")
        self.rel_threshold = float(rel_thresh)

    def set_handoff(self, mode: str = "file"):
        """How `run` hands the skeleton to K8/K9: "file" = through the reference's quality-95 JPEG (its CLI flow,
        run_preprocessing.py:137-140 -> extract_features.py:83-92; default), "memory" = the clean in-memory skeleton."""
        if mode not in ("file", "memory"):
            raise ValueError('handoff must be "file" or "memory"')
        self._ck(self._lib.fpb_set_handoff(self._h, 1 if mode == "file" else 0), "fpb_set_handoff")
        self.handoff = mode

    def _batch_roi(self, a, dtype=np.uint8):
        """[n,h,w] (or [h,w]) with h <= H, w <= W -> ([n,H,W] planes with the images at the top-left, (h, w)).
        Declares the crop size to the library so that only that region is read and written."""
        a = np.ascontiguousarray(a)
        if a.dtype != dtype:
            raise TypeError(f"expected {np.dtype(dtype)}, got {a.dtype}")
        if a.ndim == 2:
            a = a[None]
        if a.ndim != 3 or a.shape[1] > self.H or a.shape[2] > self.W:
            raise ValueError(f"expected [n,<={self.H},<={self.W}], got {a.shape}")
        n, h, w = a.shape
        if not 1 <= n <= self.max_batch:
            raise ValueError(f"batch {n} outside [1,{self.max_batch}]")
        if (h, w) == (self.H, self.W):
            self._ck(self._lib.fpb_set_stage_dims(self._h, None, 0), "fpb_set_stage_dims")
            return a, (h, w)
        wh = np.tile(np.array([w, h], np.int32), (n, 1))
        self._ck(self._lib.fpb_set_stage_dims(self._h, _ptr(wh), n), "fpb_set_stage_dims")
        pad = np.zeros((n, self.H, self.W), dtype)
        pad[:, :h, :w] = a
        return pad, (h, w)

    def _full_frames(self):
        self._ck(self._lib.fpb_set_stage_dims(self._h, None, 0), "fpb_set_stage_dims")

    def set_post_params(self, params: Optional[Dict] = None):
        if not params:
            self._ck(self._lib.fpb_set_post_params(self._h, None), "fpb_set_post_params")
            return
        p = N.PostParams(int(params.get("quality_window", 25)), float(params.get("quality_threshold", 0.15)),
                         float(params.get("coherence_threshold", 0.2)), float(params.get("min_distance", 8.0)),
                         int(params.get("margin", 30)), int(params.get("max_minutiae", 60)),
                         int(params.get("patch_radius", 15)))
        self._ck(self._lib.fpb_set_post_params(self._h, C.byref(p)), "fpb_set_post_params")

    # ------------------------------------------------------------------ EXTENSION rows G1/G2 (not in the reference)
    GABOR_DEFAULTS = dict(n_orient=16, min_period=3, max_period=25, sigma_factor=0.45, radius_factor=2.5,
                          min_amplitude=8.0, default_period=9.0)

    @classmethod
    def _gabor_params(cls, params: Optional[Dict]):
        if not params:
            return None
        d = dict(cls.GABOR_DEFAULTS); d.update(params)
        return N.GaborParams(int(d["n_orient"]), int(d["min_period"]), int(d["max_period"]), float(d["sigma_factor"]),
                             float(d["radius_factor"]), float(d["min_amplitude"]), float(d["default_period"]))

    def enable_enhanced(self, params: Optional[Dict] = None):
        """`run` additionally produces the planes "enhanced" / "gabor_response" and `freq_blocks()`."""
        g = self._gabor_params(params)
        self._ck(self._lib.fpb_enable_enhanced(self._h, C.byref(g) if g else None), "fpb_enable_enhanced")

    def disable_enhanced(self):
        self._ck(self._lib.fpb_disable_enhanced(self._h), "fpb_disable_enhanced")

    def freq_blocks(self) -> np.ndarray:
        out = np.zeros((self.last_n, self.H // 16, self.W // 16), np.float32)
        self._ck(self._lib.fpb_fetch_freq_blocks(self._h, _ptr(out), out.nbytes), "fpb_fetch_freq_blocks")
        return out

    def enhance_gabor(self, img, mask=None, params: Optional[Dict] = None):
        """orientation field of (img, mask) -> block ridge frequencies -> Gabor bank.  Returns (freq_blocks, response, enhanced)."""
        a = self._batch(img); m = self._batch(mask) if mask is not None else None
        n = a.shape[0]
        fb = np.zeros((n, self.H // 16, self.W // 16), np.float32)
        resp = np.empty((n, self.H, self.W), np.float32); enh = np.empty_like(a)
        g = self._gabor_params(params)
        self._ck(self._lib.fpb_enhance_gabor(self._h, _ptr(a), _ptr(m), n, C.byref(g) if g else None, _ptr(fb), _ptr(resp),
                                             _ptr(enh)), "fpb_enhance_gabor")
        self.last_n = n
        return fb, resp, enh

    # ------------------------------------------------------------------ whole path
    def run(self, images) -> int:
        """Host images [n,H,W] uint8 -> H2D, K1..K9, D2H of roi / counts / refined minutiae."""
        a = self._batch(images)
        self._full_frames()
        self._ck(self._lib.fpb_run_host(self._h, _ptr(a), a.shape[0]), "fpb_run_host")
        self.last_n = a.shape[0]
        return self.last_n

    def run_async(self, images) -> int:
        """`run` without the final wait: returns once the copies and kernels are enqueued; call `wait()` before reading results.
        `images` must stay alive (ideally pinned) until then.  Two pipelines used alternately keep the GPU busy while the host
        reads the previous batch's results."""
        a = self._batch(images)
        self._full_frames()
        self._ck(self._lib.fpb_run_host_async(self._h, _ptr(a), a.shape[0]), "fpb_run_host_async")
        self._pending = a                      # keeps the host buffer alive
        self.last_n = a.shape[0]
        return self.last_n

    def wait(self):
        self._ck(self._lib.fpb_wait(self._h), "fpb_wait")
        self._pending = None

    # ------------------------------------------------------------------ on-disk hand-offs (include/fpb200_io.h)
    def decode_jpeg(self, blobs, threads: int = 0) -> np.ndarray:
        """JPEG files in memory -> the handle's device input plane (Huffman on host threads, islow IDCT on the GPU;
        pixels identical to cv2.imread(IMREAD_GRAYSCALE)).  Returns the per-image status array (0 = decoded)."""
        n = len(blobs)
        if not 1 <= n <= self.max_batch:
            raise ValueError(f"batch {n} outside [1,{self.max_batch}]")
        ptrs = (C.c_char_p * n)(*blobs)
        sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
        status = np.zeros(n, np.int32)
        self._ck(self._lib.fpb_decode_jpeg_batch(self._h, ptrs, sizes, n, int(threads), _ptr(status)), "fpb_decode_jpeg_batch")
        return status

    def synth_ridge(self, seed: int, first_index: int, n: int, period: float = 0.0, noise_sigma: float = 12.0):
        """Fill the device input plane with images first_index .. first_index+n-1 of the synthetic stream `seed` (generated
        on the GPU, counter-based: BASELINE configs[3]).  Follow with `run_decoded(n)`; `fetch_input(n)` returns the pixels."""
        if not 1 <= n <= self.max_batch:
            raise ValueError(f"batch {n} outside [1,{self.max_batch}]")
        self._ck(self._lib.fpb_synth_ridge(self._h, int(seed), int(first_index), int(n), float(period), float(noise_sigma)), "fpb_synth_ridge")

    def run_input_async(self, n: int):
        """K1..K9 on the first n images of the device input plane, asynchronously (call `download()` later)."""
        self.run_device(int(self._lib.fpb_input_plane(self._h)), n)

    def fetch_input(self, n: int) -> np.ndarray:
        out = np.empty((n, self.H, self.W), np.uint8)
        self._ck(self._lib.fpb_fetch_input(self._h, _ptr(out), n), "fpb_fetch_input")
        return out

    def run_decoded(self, n: int) -> int:
        """K1..K9 on the first n images `decode_jpeg` left on the device."""
        self._full_frames()
        self._ck(self._lib.fpb_run_decoded(self._h, int(n)), "fpb_run_decoded")
        self.last_n = int(n)
        return self.last_n

    def write_json(self, paths, threads: int = 0) -> int:
        """`json.dump(minutiae(i), open(paths[i], "w"), indent=2)` for i < len(paths), natively and in parallel."""
        n = len(paths)
        arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        return self._ck(self._lib.fpb_write_minutiae_json_batch(self._h, arr, n, int(threads)), "fpb_write_minutiae_json_batch")

    def run_many(self, images):
        """Any number of images: chunks of `max_batch` through `run`; yields (global index, roi, refined minutiae)."""
        n = len(images)
        for s0 in range(0, n, self.max_batch):
            m = self.run(images[s0:s0 + self.max_batch])
            for i in range(m):
                yield s0 + i, self.roi(i), self.minutiae(i)

    def run_device(self, dev_ptr: int, n: int):
        """Device-resident images (raw pointer, n*H*W bytes); asynchronous."""
        self._full_frames()
        self._ck(self._lib.fpb_run_device(self._h, C.c_void_p(dev_ptr), int(n)), "fpb_run_device")
        self.last_n = int(n)

    def download(self):
        self._ck(self._lib.fpb_download_results(self._h), "fpb_download_results")

    def download_refined(self):
        """D2H of roi / counts / refined lists only (streaming loops: no raw lists)."""
        self._ck(self._lib.fpb_download_refined(self._h), "fpb_download_refined")

    def result_block(self, cap: int = 64):
        """(roi [n,4] int32, raw_counts [n], out_counts [n], refined [n,cap] MINUTIA_DTYPE records) of the last downloaded
        run in ONE call - the per-image accessors cost a ctypes round trip each, too slow for 10^4..10^5 images/s."""
        n = self.last_n
        roi = np.empty((n, 4), np.int32); rc = np.empty(n, np.int32); oc = np.empty(n, np.int32)
        out = np.zeros((n, cap), MINUTIA_DTYPE)
        self._ck(self._lib.fpb_result_block(self._h, _ptr(roi), _ptr(rc), _ptr(oc), _ptr(out), int(cap)), "fpb_result_block")
        return roi, rc, oc, out

    def roi(self, i: int):
        r = (C.c_int32 * 4)()
        self._ck(self._lib.fpb_result_roi(self._h, i, r), "fpb_result_roi")
        return tuple(int(v) for v in r)

    def raw_minutiae(self, i: int) -> List[Dict]:
        cnt = self._ck(self._lib.fpb_result_raw(self._h, i, None, 0), "fpb_result_raw")
        buf = np.zeros((max(cnt, 1), 3), np.int32)
        rc = self._lib.fpb_result_raw(self._h, i, _ptr(buf), cnt)
        if rc < 0:       # raw lists were not downloaded by run(): fetch them now
            self.download()
            self._ck(self._lib.fpb_result_raw(self._h, i, _ptr(buf), cnt), "fpb_result_raw")
        return [{"x": int(x), "y": int(y), "type": TYPE_NAMES[int(t)]} for x, y, t in buf[:cnt]]

    def minutiae(self, i: int) -> List[Dict]:
        cnt = self._ck(self._lib.fpb_result_minutiae(self._h, i, None, 0), "fpb_result_minutiae")
        buf = (N.Minutia * max(cnt, 1))()
        self._ck(self._lib.fpb_result_minutiae(self._h, i, buf, cnt), "fpb_result_minutiae")
        return [_minutia_dict(buf[k]) for k in range(cnt)]

    def fetch(self, name: str) -> np.ndarray:
        """Intermediate plane of the last run as [n,H,W] (crop planes: valid region [h',w'] at the origin)."""
        dt = np.float32 if name in N.F32_PLANES else np.uint8
        out = np.empty((self.last_n, self.H, self.W), dt)
        self._ck(self._lib.fpb_fetch_plane(self._h, N.PLANES[name], _ptr(out), out.nbytes), "fpb_fetch_plane")
        return out

    # ------------------------------------------------------------------ stages
    def normalize(self, img):
        a = self._batch(img); out = np.empty_like(a); self._full_frames()
        self._ck(self._lib.fpb_normalize(self._h, _ptr(a), a.shape[0], _ptr(out)), "fpb_normalize")
        return out

    def denoise(self, img, with_nlm: bool = False):
        a = self._batch(img); out = np.empty_like(a); nlm = np.empty_like(a) if with_nlm else None
        self._full_frames()
        self._ck(self._lib.fpb_denoise(self._h, _ptr(a), a.shape[0], _ptr(out), _ptr(nlm)), "fpb_denoise")
        return (out, nlm) if with_nlm else out

    def segment_bgr(self, img):
        """segment_fingerprint on colour input [n,H,W,3|4] (or [H,W,3|4]): cv2.COLOR_BGR2GRAY on the device first."""
        a = np.ascontiguousarray(img)
        if a.dtype != np.uint8:
            raise TypeError(f"expected uint8, got {a.dtype}")
        if a.ndim == 3:
            a = a[None]
        if a.ndim != 4 or a.shape[1:3] != (self.H, self.W) or a.shape[3] not in (3, 4):
            raise ValueError(f"expected [n,{self.H},{self.W},3|4], got {a.shape}")
        n = a.shape[0]
        if not 1 <= n <= self.max_batch:
            raise ValueError(f"batch {n} outside [1,{self.max_batch}]")
        seg = np.empty((n, self.H, self.W), np.uint8); mask = np.empty_like(seg)
        roi = np.zeros((n, 4), np.int32); self._full_frames()
        self._ck(self._lib.fpb_segment_bgr(self._h, _ptr(a), a.shape[3], n, _ptr(seg), _ptr(mask), _ptr(roi)), "fpb_segment_bgr")
        return seg, mask, roi

    def segment(self, img):
        a = self._batch(img); seg = np.empty_like(a); mask = np.empty_like(a)
        roi = np.zeros((a.shape[0], 4), np.int32); self._full_frames()
        self._ck(self._lib.fpb_segment(self._h, _ptr(a), a.shape[0], _ptr(seg), _ptr(mask), _ptr(roi)), "fpb_segment")
        return seg, mask, roi

    # The stages below work on the data-dependent CROP of segment_fingerprint: they accept [n,h,w] with h <= H, w <= W
    # (one handle serves a whole bucket of crop sizes, `pipeline_for`) and return arrays of the caller's size.
    def binarize(self, img):
        a, (h, w) = self._batch_roi(img); out = np.empty_like(a)
        self._ck(self._lib.fpb_binarize(self._h, _ptr(a), a.shape[0], _ptr(out)), "fpb_binarize")
        return out[:, :h, :w]

    def orientation(self, img, mask=None, block_size: int = 16, smooth_sigma: float = 3.0,
                    invert_if_needed: bool = True, smooth_orientation_sigma: float = 3.0):
        """compute_orientation_map (orientation.py:9-85); the keyword defaults are the hot path's values (fpb_orientation),
        anything else goes through fpb_orientation_ex."""
        is_f32 = np.asarray(img).dtype == np.float32              # non-uint8 input (orientation.py:21-24): fpb_orientation_f32
        a, (h, w) = self._batch_roi(img, np.float32 if is_f32 else np.uint8)
        m = self._batch_roi(mask)[0] if mask is not None else None
        n = a.shape[0]
        bs = int(block_size)
        if bs < 1:
            raise ValueError("block_size must be >= 1")
        blocks = np.zeros((n, self.H // bs, self.W // bs), np.float32)
        oimg = np.empty((n, self.H, self.W), np.float32); rel = np.empty_like(oimg)
        if is_f32:
            self._ck(self._lib.fpb_orientation_f32(self._h, _ptr(a), _ptr(m), n, bs, float(smooth_sigma),
                                                   int(bool(invert_if_needed)), float(smooth_orientation_sigma),
                                                   _ptr(blocks), _ptr(oimg), _ptr(rel)), "fpb_orientation_f32")
        elif (bs, float(smooth_sigma), bool(invert_if_needed), float(smooth_orientation_sigma)) == (16, 3.0, True, 3.0):
            self._ck(self._lib.fpb_orientation(self._h, _ptr(a), _ptr(m), n, _ptr(blocks), _ptr(oimg), _ptr(rel)),
                     "fpb_orientation")
        else:
            self._ck(self._lib.fpb_orientation_ex(self._h, _ptr(a), _ptr(m), n, bs, float(smooth_sigma),
                                                  int(bool(invert_if_needed)), float(smooth_orientation_sigma),
                                                  _ptr(blocks), _ptr(oimg), _ptr(rel)), "fpb_orientation_ex")
        return blocks[:, :h // bs, :w // bs], oimg[:, :h, :w], rel[:, :h, :w]

    def smooth(self, binary, sigma: float = 1.4, diffusion_iter: int = 3, contrast_boost: float = 1.25,
               force_unfused: bool = False):
        """smooth_fingerprint_skeleton (fingerprint_preprocess.py:141-159); non-default keyword values (or
        `force_unfused`, for tests) run the unfused kernel sequence through fpb_smooth_ex."""
        a, (h, w) = self._batch_roi(binary); out = np.empty_like(a)
        if not force_unfused and (float(sigma), int(diffusion_iter), float(contrast_boost)) == (1.4, 3, 1.25):
            self._ck(self._lib.fpb_smooth(self._h, _ptr(a), a.shape[0], _ptr(out)), "fpb_smooth")
        else:
            self._ck(self._lib.fpb_smooth_ex(self._h, _ptr(a), a.shape[0], float(sigma), int(diffusion_iter),
                                             float(contrast_boost), _ptr(out)), "fpb_smooth_ex")
        return out[:, :h, :w]

    def thin(self, binary_smooth, reliability, with_gate: bool = False, rel_thresh: float = 0.1):
        if float(rel_thresh) != self.rel_threshold:
            self.set_rel_threshold(rel_thresh)
        a, (h, w) = self._batch_roi(binary_smooth); r = self._batch_roi(reliability, np.float32)[0]
        out = np.empty_like(a); gate = np.empty_like(a) if with_gate else None
        self._ck(self._lib.fpb_thin(self._h, _ptr(a), _ptr(r), a.shape[0], _ptr(out), _ptr(gate)), "fpb_thin")
        return (out[:, :h, :w], gate[:, :h, :w]) if with_gate else out[:, :h, :w]

    def skeletonize(self, gate):
        a, (h, w) = self._batch_roi(gate); out = np.empty_like(a)
        self._ck(self._lib.fpb_skeletonize(self._h, _ptr(a), a.shape[0], _ptr(out)), "fpb_skeletonize")
        return out[:, :h, :w]

    def jpeg_roundtrip(self, img):
        """= cv2.imread(cv2.imwrite(img as .jpg, quality 95), IMREAD_GRAYSCALE): the reference's skeleton hand-off."""
        a, (h, w) = self._batch_roi(img); out = np.empty_like(a)
        self._ck(self._lib.fpb_jpeg_roundtrip(self._h, _ptr(a), a.shape[0], _ptr(out)), "fpb_jpeg_roundtrip")
        return out[:, :h, :w]

    def extract_minutiae(self, skel, cap: Optional[int] = None) -> List[List[Dict]]:
        a, _ = self._batch_roi(skel); n = a.shape[0]
        cap = self.raw_capacity if cap is None else int(cap)
        counts = np.zeros(n, np.int32); xyt = np.zeros((n, cap, 3), np.int32)
        self._ck(self._lib.fpb_extract_minutiae(self._h, _ptr(a), n, _ptr(counts), _ptr(xyt), cap), "fpb_extract_minutiae")
        if counts.max(initial=0) > cap:
            raise N.FpbError(f"{int(counts.max())} raw minutiae exceed the list capacity {cap}")
        return [[{"x": int(x), "y": int(y), "type": TYPE_NAMES[int(t)]} for x, y, t in xyt[b, :counts[b]]]
                for b in range(n)]

    def postprocess(self, skel, raw_lists: List[List[Dict]], cap_out: int = 128, gray=None) -> List[List[Dict]]:
        """postprocess_minutiae; `gray` (same shape as `skel`, uint8) is the image the orientation / coherence maps are
        computed from - None: the skeleton itself, the reference's call site (extract_features.py:92)."""
        a, _ = self._batch_roi(skel); n = a.shape[0]
        g = None
        if gray is not None:
            if np.shape(gray) != np.shape(skel):
                raise ValueError(f"gray {np.shape(gray)} and skel {np.shape(skel)} differ in shape")
            g = self._batch_roi(gray)[0]
        cap = max(1, max(len(r) for r in raw_lists))
        counts = np.array([len(r) for r in raw_lists], np.int32)
        xyt = np.zeros((n, cap, 3), np.int32)
        for b, lst in enumerate(raw_lists):
            for k, m in enumerate(lst):
                xyt[b, k] = (int(m["x"]), int(m["y"]), 0 if m["type"] == "ending" else 1)
        out_counts = np.zeros(n, np.int32)
        out = (N.Minutia * (n * cap_out))()
        if g is None:
            self._ck(self._lib.fpb_postprocess(self._h, _ptr(a), n, _ptr(counts), _ptr(xyt), cap, _ptr(out_counts),
                                               out, cap_out), "fpb_postprocess")
        else:
            self._ck(self._lib.fpb_postprocess_gray(self._h, _ptr(a), _ptr(g), n, _ptr(counts), _ptr(xyt), cap,
                                                    _ptr(out_counts), out, cap_out), "fpb_postprocess_gray")
        return [[_minutia_dict(out[b * cap_out + k]) for k in range(min(int(out_counts[b]), cap_out))] for b in range(n)]


    def nms_adaptive(self, xy, quality, density, base_dist: float) -> np.ndarray:
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2); n = len(xy)
        q = np.ascontiguousarray(quality, np.float64); d = np.ascontiguousarray(density, np.float32)
        keep = np.zeros(n, np.uint8)
        self._ck(self._lib.fpb_nms_adaptive(self._h, n, _ptr(xy), _ptr(q), _ptr(d), float(base_dist), _ptr(keep)), "fpb_nms_adaptive")
        return keep.astype(bool)

    def remove_redundant(self, xy, quality, orientation, density, base_radius: float, angle_thresh: float) -> np.ndarray:
        xy = np.ascontiguousarray(xy, np.int32).reshape(-1, 2); n = len(xy)
        q = np.ascontiguousarray(quality, np.float64); o = np.ascontiguousarray(orientation, np.float64)
        d = np.ascontiguousarray(density, np.float32)
        keep = np.zeros(n, np.uint8)
        self._ck(self._lib.fpb_remove_redundant(self._h, n, _ptr(xy), _ptr(q), _ptr(o), _ptr(d), float(base_radius),
                                                float(angle_thresh), _ptr(keep)), "fpb_remove_redundant")
        return keep.astype(bool)


def _minutia_dict(m) -> Dict:
    return {"x": int(m.x), "y": int(m.y), "type": TYPE_NAMES[int(m.type)], "orientation": float(m.orientation),
            "quality": float(m.quality), "coherence": float(m.coherence),
            "angular_stability": float(m.angular_stability)}


# ---------------------------------------------------------------------- thinning table override
def load_thin_table(path: str) -> np.ndarray:
    """256 entries 0..3 in scikit-image's neighbour coding: raw 256-byte file, .npy, or text (whitespace / commas)."""
    if path.endswith(".npy"):
        t = np.load(path)
    else:
        blob = open(path, "rb").read()
        if len(blob) == 256:
            t = np.frombuffer(blob, np.uint8)
        else:
            t = np.array([int(v) for v in blob.decode().replace(",", " ").replace("[", " ").replace("]", " ").split()])
    t = np.asarray(t).reshape(-1)
    if t.size != 256 or t.min() < 0 or t.max() > 3:
        raise ValueError(f"{path}: a thinning table has 256 entries in 0..3")
    return t.astype(np.uint8)


def thin_table_from_env() -> Optional[np.ndarray]:
    """FPB200_THIN_TABLE=<file>: the 256-entry deletion table every handle of this process uses instead of the built-in
    Zhang-Suen one (e.g. scikit-image's literal `_fast_skeletonize` table dumped on a machine that has it)."""
    path = os.environ.get("FPB200_THIN_TABLE")
    return load_thin_table(path) if path else None


# ---------------------------------------------------------------------- per-thread handle cache
_tls = threading.local()
BUCKET = 64                      # crop sizes are rounded up to multiples of this


def _bucket(v: int) -> int:
    return max(BUCKET, (int(v) + BUCKET - 1) // BUCKET * BUCKET)


def pipeline_for(height: int, width: int, max_batch: int = 1, device: int = 0, exact: bool = False) -> FingerprintPipeline:
    """Cached handle for the calling thread (the reference's functions are called concurrently from
    ThreadPoolExecutor workers - run_preprocessing.py:154 - so handles are never shared across threads).

    `exact=False` (the crop stages: binarize ... postprocess): the handle is sized for the 64-pixel BUCKET the shape
    falls in and the per-image size travels as a ROI, so the data-dependent crops of the reference's per-file flow
    (extract_features.process_image on <base>_skeleton.jpg) share a handful of handles instead of creating a workspace
    per crop size.  `exact=True` (normalize / denoise / segment / the fused run) needs whole frames of that size."""
    cache = getattr(_tls, "cache", None)
    if cache is None:
        cache = _tls.cache = {}
    H, W = int(height), int(width)
    if not exact:
        Hb, Wb = _bucket(H), _bucket(W)
        if Hb <= 16383 and Wb <= 16383 and ((Wb + 31) // 32) * Hb <= 32768:   # the library's shape limits (fpb_create)
            H, W = Hb, Wb
    key = (H, W, int(max_batch), int(device))
    p = cache.get(key)
    if p is None:
        if len(cache) >= 16:                     # bound the cache
            cache.pop(next(iter(cache))).close()
        p = cache[key] = FingerprintPipeline(H, W, max_batch, device)
    return p


def handle_cache_size() -> int:
    """Number of live handles of the calling thread (tests: crops of many sizes must share a few)."""
    return len(getattr(_tls, "cache", None) or {})
