"""ctypes binding of libfpb200.so (include/fpb200.h) - the only compute backend.

There is deliberately NO fallback: if the CUDA library has not been built, or no CUDA
device is usable, importing / creating a handle raises.  Build it with
`python -c "import __graft_entry__ as g; g.build()"` (or `make -C .../csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfpb200.so")


class FpbError(RuntimeError):
    pass


class PostParams(C.Structure):
    _fields_ = [("quality_window", C.c_int), ("quality_threshold", C.c_double),
                ("coherence_threshold", C.c_double), ("min_distance", C.c_double),
                ("margin", C.c_int), ("max_minutiae", C.c_int), ("patch_radius", C.c_int)]


class Minutia(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("type", C.c_int32), ("_pad", C.c_int32),
                ("orientation", C.c_double), ("quality", C.c_double), ("coherence", C.c_double),
                ("angular_stability", C.c_double)]


class GaborParams(C.Structure):            # fpb_gabor_params (EXTENSION rows G1/G2)
    _fields_ = [("n_orient", C.c_int), ("min_period", C.c_int), ("max_period", C.c_int), ("sigma_factor", C.c_double),
                ("radius_factor", C.c_double), ("min_amplitude", C.c_double), ("default_period", C.c_double)]


class MatchParams(C.Structure):            # fpb_match_params (include/fpb200_match.h)
    _fields_ = [("dist_thresh", C.c_double), ("orient_thresh_deg", C.c_double), ("use_type", C.c_int),
                ("ransac_iter", C.c_int), ("min_inliers", C.c_int), ("stop_inlier_ratio", C.c_double),
                ("cross_check", C.c_int)]


class MatchResult(C.Structure):            # fpb_match_result
    _fields_ = [("final_score", C.c_double), ("inlier_ratio", C.c_double), ("theta", C.c_double),
                ("tx", C.c_double), ("ty", C.c_double), ("n_matches", C.c_int32), ("best_iter", C.c_int32)]


PLANES = {"normalized": 0, "denoised": 1, "segmented": 2, "mask": 3, "binary": 4, "binary_smooth": 5,
          "skeleton": 6, "orient_img": 7, "reliability": 8, "gate": 9, "nlm": 10,
          "skel_orient_img": 11, "skel_coherence": 12, "density": 13, "enhanced": 14, "gabor_response": 15,
          "skeleton_file": 16}
F32_PLANES = {"orient_img", "reliability", "skel_orient_img", "skel_coherence", "density", "gabor_response"}

# name -> (restype, argtypes); every symbol include/fpb200.h declares
_vp, _i, _sz = C.c_void_p, C.c_int, C.c_size_t
SIGNATURES = {
    "fpb_abi_version": (_i, []),
    "fpb_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _vp]),
    "fpb_destroy": (None, [_vp]),
    "fpb_last_error": (C.c_char_p, [_vp]),
    "fpb_sync": (_i, [_vp]),
    "fpb_set_thin_table": (_i, [_vp, _vp]),
    "fpb_set_handoff": (_i, [_vp, _i]),
    "fpb_set_rel_threshold": (_i, [_vp, C.c_double]),
    "fpb_set_stage_dims": (_i, [_vp, _vp, _i]),
    "fpb_raw_capacity": (_i, [_vp]),
    "fpb_set_post_params": (_i, [_vp, C.POINTER(PostParams)]),
    "fpb_run_device": (_i, [_vp, _vp, _i]),
    "fpb_run_host": (_i, [_vp, _vp, _i]),
    "fpb_run_host_async": (_i, [_vp, _vp, _i]),
    "fpb_wait": (_i, [_vp]),
    "fpb_download_results": (_i, [_vp]),
    "fpb_download_refined": (_i, [_vp]),
    "fpb_result_block": (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    "fpb_result_roi": (_i, [_vp, _i, _vp]),
    "fpb_result_raw": (_i, [_vp, _i, _vp, _i]),
    "fpb_result_minutiae": (_i, [_vp, _i, _vp, _i]),
    "fpb_fetch_plane": (_i, [_vp, _i, _vp, _sz]),
    "fpb_launch_count": (C.c_longlong, [_vp]),
    "fpb_set_profiling": (_i, [_vp, _i]),
    "fpb_stage_times": (_i, [_vp, _vp, _i]),
    "fpb_kernel_times": (_i, [_vp, _i, C.c_char_p, _i]),
    "fpb_normalize": (_i, [_vp, _vp, _i, _vp]),
    "fpb_denoise": (_i, [_vp, _vp, _i, _vp, _vp]),
    "fpb_segment": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "fpb_segment_bgr": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    "fpb_binarize": (_i, [_vp, _vp, _i, _vp]),
    "fpb_orientation": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "fpb_orientation_ex": (_i, [_vp, _vp, _vp, _i, _i, C.c_double, _i, C.c_double, _vp, _vp, _vp]),
    "fpb_orientation_f32": (_i, [_vp, _vp, _vp, _i, _i, C.c_double, _i, C.c_double, _vp, _vp, _vp]),
    "fpb_smooth": (_i, [_vp, _vp, _i, _vp]),
    "fpb_smooth_ex": (_i, [_vp, _vp, _i, C.c_double, _i, C.c_double, _vp]),
    "fpb_thin": (_i, [_vp, _vp, _vp, _i, _vp, _vp]),
    "fpb_skeletonize": (_i, [_vp, _vp, _i, _vp]),
    "fpb_extract_minutiae": (_i, [_vp, _vp, _i, _vp, _vp, _i]),
    "fpb_postprocess": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i]),
    "fpb_postprocess_gray": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i]),
    "fpb_nms_adaptive": (_i, [_vp, _i, _vp, _vp, _vp, C.c_double, _vp]),
    "fpb_remove_redundant": (_i, [_vp, _i, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp]),
    "fpb_enable_enhanced": (_i, [_vp, C.POINTER(GaborParams)]),
    "fpb_disable_enhanced": (_i, [_vp]),
    "fpb_enhance_gabor": (_i, [_vp, _vp, _vp, _i, C.POINTER(GaborParams), _vp, _vp, _vp]),
    "fpb_fetch_freq_blocks": (_i, [_vp, _vp, _sz]),
    # include/fpb200_io.h
    "fpb_jpeg_roundtrip": (_i, [_vp, _vp, _i, _vp]),
    "fpb_synth_ridge": (_i, [_vp, C.c_uint64, C.c_uint64, _i, C.c_double, C.c_double]),
    "fpb_jpeg_info": (_i, [_vp, _sz, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "fpb_jpeg_coefficients": (_i, [_vp, _sz, _i, _i, _vp, _vp]),
    "fpb_decode_jpeg_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "fpb_fetch_input": (_i, [_vp, _vp, _i]),
    "fpb_run_decoded": (_i, [_vp, _i]),
    "fpb_input_plane": (_vp, [_vp]),
    "fpb_minutiae_json": (C.c_longlong, [_vp, _i, _vp, _sz]),
    "fpb_write_minutiae_json_batch": (_i, [_vp, _vp, _i, _i]),
    # include/fpb200_unet.h
    "fpb_unet_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "fpb_unet_destroy": (None, [_vp]),
    "fpb_unet_last_error": (C.c_char_p, [_vp]),
    "fpb_unet_conv_shape": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "fpb_unet_set_conv": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, C.c_double]),
    "fpb_unet_set_final": (_i, [_vp, _vp, _vp]),
    "fpb_unet_forward": (_i, [_vp, _vp, _i, _vp]),
    "fpb_unet_launches": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    # include/fpb200_match.h
    "fpb_match_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "fpb_match_destroy": (None, [_vp]),
    "fpb_match_last_error": (C.c_char_p, [_vp]),
    "fpb_match_set_templates": (_i, [_vp, _vp, _vp, _i]),
    "fpb_match_pairs": (_i, [_vp, _vp, _i, C.POINTER(MatchParams), _vp, _vp, _vp]),
    "fpb_match_upload_pairs": (_i, [_vp, _vp, _i]),
    "fpb_match_run_device": (_i, [_vp, C.POINTER(MatchParams)]),
    "fpb_match_download": (_i, [_vp, _vp, _vp, _vp]),
    "fpb_match_sync": (_i, [_vp]),
    "fpb_match_stream": (_vp, [_vp]),
    "fpb_match_launch_count": (C.c_longlong, [_vp]),
    "fpb_match_seed_uniforms": (_i, [C.c_uint64, _i, _vp]),
}

_lib = None


def load():
    """Load libfpb200.so (once).  Raises FpbError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FpbError(f"{LIB_PATH} is missing: build the CUDA library first "
                       "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
