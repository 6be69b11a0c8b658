from oracle.skimage_compat import (  # noqa: F401
    remove_small_objects, remove_small_holes, reconstruction, skeletonize)
