"""The oracle's OpenCV-primitive restatements (oracle/cvprim.c) against golden vectors
produced by cv2 4.13.0 (scripts/make_golden.py), and against live cv2 when importable."""
import ctypes

import numpy as np
import pytest


@pytest.fixture(scope="module")
def L(oracle_mod):
    lib = oracle_mod.Oracle.lib()
    lib.cvp_fast_atan2.restype = ctypes.c_float
    lib.cvp_fast_atan2.argtypes = [ctypes.c_float, ctypes.c_float]
    return lib


class Corner(ctypes.Structure):
    _fields_ = [("x", ctypes.c_int), ("y", ctypes.c_int), ("s", ctypes.c_int)]


def _resize(L, src, dw, dh):
    out = np.zeros((dh, dw), np.uint8)
    L.cvp_resize_linear_u8(ctypes.c_void_p(src.ctypes.data), src.shape[1], src.shape[0], src.strides[0],
                           ctypes.c_void_p(out.ctypes.data), dw, dh, out.strides[0])
    return out


def _blur(L, src):
    out = np.zeros_like(src)
    L.cvp_gaussian7x7_s2_u8(ctypes.c_void_p(src.ctypes.data), src.shape[1], src.shape[0], src.strides[0],
                            ctypes.c_void_p(out.ctypes.data), out.strides[0])
    return out


def _fast(L, img, th, nms):
    buf = (Corner * img.size)()
    n = L.cvp_fast9_16(ctypes.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], th, nms, buf, img.size)
    return np.array([(buf[i].x, buf[i].y, buf[i].s) for i in range(n)], np.int32).reshape(-1, 3)


def test_resize_golden(L, golden_prims):
    src = golden_prims["resize_src"]
    for k in range(5):
        ref = golden_prims["resize_dst_%d" % k]
        assert np.array_equal(_resize(L, src, ref.shape[1], ref.shape[0]), ref), k


def test_blur_golden(L, golden_prims):
    assert np.array_equal(_blur(L, golden_prims["blur_src"]), golden_prims["blur_dst"])
    assert np.array_equal(_blur(L, golden_prims["blur_small_src"]), golden_prims["blur_small_dst"])


def test_border_golden(L, golden_prims):
    src = golden_prims["blur_src"]
    ref = golden_prims["border_dst"]
    out = np.zeros_like(ref)
    L.cvp_border_reflect101_u8(ctypes.c_void_p(src.ctypes.data), src.shape[1], src.shape[0], src.strides[0],
                               ctypes.c_void_p(out.ctypes.data), out.strides[0], 19, 19, 19, 19)
    assert np.array_equal(out, ref)


def test_fast_golden(L, golden_prims):
    img = golden_prims["fast_src"]
    for th in (20, 7):
        ref = golden_prims["fast_t%d_nms1" % th]
        assert np.array_equal(_fast(L, img, th, 1), ref), th          # positions, ORDER and scores
        ref0 = golden_prims["fast_t%d_nms0" % th]
        assert np.array_equal(_fast(L, img, th, 0)[:, :2], ref0[:, :2]), th


def test_fast_atan2_golden(L, golden_prims):
    out = np.array([L.cvp_fast_atan2(float(y), float(x)) for y, x in zip(golden_prims["atan2_y"], golden_prims["atan2_x"])], np.float32)
    assert np.array_equal(out.view(np.uint32), golden_prims["atan2_deg"].view(np.uint32))


def test_rounding_helpers(L):
    L.cvp_round_f.argtypes = [ctypes.c_float]
    L.cvp_round_d.argtypes = [ctypes.c_double]
    assert [L.cvp_round_f(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 2.4999)] == [0, 2, 2, 0, -2, 2]   # half to even
    assert L.cvp_round_d(1241.5) == 1242 and L.cvp_round_d(1242.5) == 1242


def test_primitives_against_live_cv2(L):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for _ in range(12):
        h, w = (int(v) for v in rng.integers(24, 200, 2))
        dh, dw = (int(v) for v in rng.integers(12, 200, 2))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(_resize(L, img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
        assert np.array_equal(_blur(L, img), cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        det = cv2.FastFeatureDetector_create(threshold=12, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        ref = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)], np.int32).reshape(-1, 3)
        assert np.array_equal(_fast(L, img, 12, 1), ref)


def _gray(L, img, rgb_order):
    h, w, c = img.shape
    out = np.zeros((h, w), np.uint8)
    L.cvp_cvt_gray_u8(ctypes.c_void_p(img.ctypes.data), w, h, img.strides[0], c, int(rgb_order), ctypes.c_void_p(out.ctypes.data), out.strides[0])
    return out


def test_gray_conversion_against_live_cv2(L):
    """cvtColor(…, CV_RGB2GRAY / BGR2GRAY / RGBA2GRAY / BGRA2GRAY) as Tracking::GrabImage* calls it (src/Tracking.cc:459-472)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    for c in (3, 4):
        img = rng.integers(0, 256, (57, 131, c), dtype=np.uint8)
        img[0, :8] = [[0] * c, [255] * c, [255, 0, 0, 9][:c], [0, 255, 0, 9][:c], [0, 0, 255, 9][:c], [1, 1, 1, 1][:c], [254, 255, 253, 0][:c], [128] * c]
        codes = {(3, 1): cv2.COLOR_RGB2GRAY, (3, 0): cv2.COLOR_BGR2GRAY, (4, 1): cv2.COLOR_RGBA2GRAY, (4, 0): cv2.COLOR_BGRA2GRAY}
        for rgb in (1, 0):
            assert np.array_equal(_gray(L, img, rgb), cv2.cvtColor(img, codes[(c, rgb)]))


def test_vectorised_primitives_equal_their_scalar_forms(L):
    """The AVX2 / vectoriser-friendly forms of FAST and the blur (the ones the reference-compiled CPU baseline runs on) against the
    plain scalar restatements, on sizes around the 32-pixel chunk and the cell sizes the reference calls cv::FAST with (36 .. 38)."""
    rng = np.random.default_rng(11)
    from multimot_track_b200.synth import value_noise_frame
    for (h, w) in ((7, 7), (8, 9), (36, 36), (36, 37), (38, 38), (37, 39), (20, 70), (41, 67), (64, 99), (90, 131)):
        for kind in range(3):
            if kind == 0:
                img = rng.integers(0, 256, (h, w), dtype=np.uint8)
            elif kind == 1:
                img = value_noise_frame(h + w, max(h, 24), max(w, 24))[:h, :w].copy()
            else:
                img = (rng.integers(0, 2, (h, w)) * 255).astype(np.uint8)            # saturating differences
            big = np.zeros((h, w + 13), np.uint8); big[:, :w] = img
            view = big[:, :w]                                                        # stride != width
            for th in (1, 7, 20, 100, 254):
                for nms in (1, 0):
                    buf = (Corner * img.size)(); buf2 = (Corner * img.size)()
                    n = L.cvp_fast9_16(ctypes.c_void_p(view.ctypes.data), w, h, view.strides[0], th, nms, buf, img.size)
                    n2 = L.cvp_fast9_16_scalar(ctypes.c_void_p(view.ctypes.data), w, h, view.strides[0], th, nms, buf2, img.size)
                    assert n == n2, (h, w, kind, th, nms)
                    assert all((buf[i].x, buf[i].y, buf[i].s) == (buf2[i].x, buf2[i].y, buf2[i].s) for i in range(n)), (h, w, kind, th, nms)
            a = np.zeros_like(img); b = np.zeros_like(img)
            L.cvp_gaussian7x7_s2_u8(ctypes.c_void_p(view.ctypes.data), w, h, view.strides[0], ctypes.c_void_p(a.ctypes.data), a.strides[0])
            L.cvp_gaussian7x7_s2_u8_scalar(ctypes.c_void_p(view.ctypes.data), w, h, view.strides[0], ctypes.c_void_p(b.ctypes.data), b.strides[0])
            assert np.array_equal(a, b), (h, w, kind)
