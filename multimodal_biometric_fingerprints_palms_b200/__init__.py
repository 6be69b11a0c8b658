"""B200-native fingerprint enhance -> minutiae hot path.

Drop-in for `src.preprocessing` / `src.features` of
GiovanniIacuzzo/multimodal_biometric_fingerprints_palms (same function names, arguments,
return containers and error behaviour - SURVEY.md section 8(b)); the arithmetic runs in the
hand-written sm_100a kernels of `libfpb200.so` (C ABI: include/fpb200.h).  No CPU fallback.

    from multimodal_biometric_fingerprints_palms_b200.preprocessing.fingerprint_preprocess import preprocess_fingerprint
    from multimodal_biometric_fingerprints_palms_b200.features.extract_features import extract_minutiae
    from multimodal_biometric_fingerprints_palms_b200.features.post_processing import postprocess_minutiae
    from multimodal_biometric_fingerprints_palms_b200 import FingerprintPipeline      # batched API
"""
from .pipeline import FingerprintPipeline, pipeline_for  # noqa: F401
from ._native import FpbError  # noqa: F401

__version__ = "0.1.0"
