"""Import shim: routes the reference's `skimage` imports to oracle.skimage_compat
(scikit-image is absent from this image).  Used only by oracle/make_golden.py."""
