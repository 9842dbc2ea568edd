"""sfm_gms_b200 — B200-native BF-Hamming + GMS matching stage (drop-in for the reference's
``BFMatcher::match`` + ``cv::xfeatures2d::matchGMS`` call pair, FeatureMatchUtil.cpp:66-69).

The compute path is the in-tree CUDA library ``libsfmgms.so`` (C ABI: include/sfmgms.h).  There is no
CPU or PyTorch fallback: importing works anywhere, but creating a :class:`Context` without the built
library or without a B200 raises.
"""
from .api import (  # noqa: F401
    NORM_HAMMING,
    NORM_L2,
    ORB,
    ORB_create,
    BFMatcher,
    Context,
    DMatch,
    KeyPoint,
    SfmGmsError,
    bruteForceMatch,
    default_context,
    gms_matcher,
    load_library,
    matchGMS,
)

__all__ = ["NORM_HAMMING", "NORM_L2", "ORB", "ORB_create", "BFMatcher", "Context", "DMatch", "KeyPoint", "SfmGmsError", "bruteForceMatch", "default_context", "gms_matcher",
           "load_library", "matchGMS"]
