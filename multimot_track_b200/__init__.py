"""B200-native ORB feature front end for MultMotTracking's ORB-SLAM2-derived tracker.

Host-side mirror of the reference's `ORBextractor` / `ORBmatcher` surface
(include/ORBextractor.h:45-111, include/ORBmatcher.h:37-109 in the reference tree)
over the C ABI of include/orbx.h (liborbx.so, hand-written sm_100a CUDA kernels).
There is no CPU fallback: importing works anywhere, computing needs a CUDA device.
"""
from ._lib import OrbxError, lib_path, load_library  # noqa: F401
from .extractor import COMPACT_KEYPOINT_DTYPE, KEYPOINT_DTYPE, ORBextractor  # noqa: F401
from .matcher import ORBmatcher  # noqa: F401
from .vocabulary import ORBVocabulary  # noqa: F401
from .pool import ExtractorPool  # noqa: F401

__all__ = ["ORBextractor", "ORBmatcher", "ORBVocabulary", "ExtractorPool", "KEYPOINT_DTYPE", "COMPACT_KEYPOINT_DTYPE", "OrbxError", "load_library", "lib_path"]
