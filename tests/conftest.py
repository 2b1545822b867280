import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _read_png_gray(path):
    try:
        import cv2
        return cv2.imread(path, cv2.IMREAD_UNCHANGED)
    except ImportError:                       # tiny fallback so fixtures load without cv2
        from PIL import Image
        return np.array(Image.open(path))


@pytest.fixture(scope="session")
def golden_kitti():
    return np.load(os.path.join(GOLDEN, "golden_kitti.npz"))


@pytest.fixture(scope="session")
def golden_prims():
    return np.load(os.path.join(GOLDEN, "golden_prims.npz"))


@pytest.fixture(scope="session")
def golden_synth():
    return np.load(os.path.join(GOLDEN, "golden_synth.npz"))


@pytest.fixture(scope="session")
def kitti_frames(golden_kitti):
    frames = [_read_png_gray(os.path.join(GOLDEN, "kitti_gray_%06d.png" % i)) for i in range(5)]
    for i, f in enumerate(frames):
        assert f.dtype == np.uint8 and f.shape == (375, 1242)
        assert zlib.crc32(f.tobytes()) == int(golden_kitti["gray_crc_%d" % i])
    return frames


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as om
    om.build(ref=True)          # compiles the port; the reference-based libs only when /root/reference exists
    return om


def crc32(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def kps_equal_exact(a, b, fields=("x", "y", "size", "response", "octave", "class_id")):
    if len(a) != len(b):
        return False
    return all(np.array_equal(a[f], b[f]) for f in fields)


def angle_diff_rad(a_deg, b_deg):
    d = np.abs(a_deg.astype(np.float64) - b_deg.astype(np.float64))
    d = np.minimum(d, 360.0 - d)
    return np.deg2rad(d)


def desc_bit_mismatch(a, b):
    assert a.shape == b.shape
    return int(np.unpackbits(a ^ b).sum()), a.size * 8
