from oracle.skimage_compat import threshold_otsu  # noqa: F401
