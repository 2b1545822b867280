"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/orbx.h
declares, refuses to compute without a GPU (no CPU fallback), and its host-only
planning arithmetic (constructor tables, level sizes, cell grid) equals the oracle's."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "multimot_track_b200", "liborbx.so")):
        ge.build()
    from multimot_track_b200 import load_library
    return load_library()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "orbx.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    from multimot_track_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.orbx_version()


def test_keypoint_layout_is_cv_keypoint():
    from multimot_track_b200 import KEYPOINT_DTYPE
    assert KEYPOINT_DTYPE.itemsize == 28
    assert [KEYPOINT_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from multimot_track_b200 import ORBextractor, OrbxError
    with pytest.raises(OrbxError) as e:
        ORBextractor(2000, 1.2, 8, 20, 7)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_bad_arguments(lib):
    from multimot_track_b200 import _lib
    h = ctypes.c_void_p()
    for cfg in (_lib.OrbxConfig(2000, 1.2, 0, 20, 7, 0, 0, 0, -1), _lib.OrbxConfig(2000, 1.2, 17, 20, 7, 0, 0, 0, -1),
                _lib.OrbxConfig(0, 1.2, 8, 20, 7, 0, 0, 0, -1), _lib.OrbxConfig(2000, 0.5, 8, 20, 7, 0, 0, 0, -1)):
        assert lib.orbx_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
        assert lib.orbx_last_error(None)
    assert lib.orbx_create(None, ctypes.byref(h)) == -1
    lib.orbx_destroy(None)                                       # no-op
    assert lib.orbx_sync(None) == -1 and lib.orbx_launch_count(None) == 0


def test_hamming256_host_scalar(lib):
    from multimot_track_b200 import ORBmatcher
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == int(np.unpackbits(a ^ b).sum())
    assert ORBmatcher.DescriptorDistance(bytes(32), bytes([255] * 32)) == 256
    assert (ORBmatcher.TH_LOW, ORBmatcher.TH_HIGH, ORBmatcher.HISTO_LENGTH) == (50, 100, 30)


@pytest.mark.parametrize("shape,params", [((375, 1242), (2000, 1.2, 8, 20, 7)), ((375, 1242), (4000, 1.2, 8, 20, 7)),
                                          ((1080, 1920), (5000, 1.2, 8, 20, 7)), ((2160, 3840), (10000, 1.2, 12, 20, 7)),
                                          ((480, 640), (1000, 1.2, 8, 20, 7)), ((300, 700), (500, 1.5, 4, 25, 10))])
def test_plan_equals_oracle_geometry(lib, oracle_mod, shape, params):
    """Host tables of the product (float arithmetic of src/ORBextractor.cc:410-470, :771-806, :1116) vs the oracle."""
    from multimot_track_b200._lib import make_plan
    h, w = shape
    plan = make_plan(*params, w, h)
    o = oracle_mod.Oracle(*params)
    t = o.tables()
    L = params[2]
    assert plan.nlevels == L
    for name, key in (("scale", "scale"), ("inv_scale", "inv_scale"), ("sigma2", "sigma2"), ("inv_sigma2", "inv_sigma2")):
        got = np.array(list(getattr(plan, name))[:L], np.float32)
        assert np.array_equal(got.view(np.uint32), t[key].view(np.uint32)), name
    assert list(plan.nfeatures_per_level)[:L] == t["nfeat"].tolist()
    assert list(plan.umax) == t["umax"].tolist()
    o(np.zeros((h, w), np.uint8))                      # a flat image: the oracle still walks the pyramid and the cells
    for l in range(L):
        assert (plan.level_width[l], plan.level_height[l]) == o.level_size(l)
        _, visited = o.level_min_cells(l)
        # the oracle counts every visited cell, the plan only rows that can hold a corner (>= 7 pixel rows)
        assert plan.cell_cols[l] * plan.cell_rows[l] <= visited <= plan.cell_cols[l] * (plan.cell_rows[l] + 1)
        assert plan.keypoint_size[l] == float(int(31 * t["scale"][l]))
    assert plan.max_keypoints >= params[0] + 3 * L


def test_plan_rejects_shapes_the_reference_cannot_process(lib):
    from multimot_track_b200._lib import make_plan
    from multimot_track_b200 import OrbxError
    with pytest.raises(OrbxError) as e:
        make_plan(1000, 1.2, 8, 20, 7, 200, 200)           # level 7 would be 56x56: nCols == 0 in the reference
    assert e.value.code == -5
    with pytest.raises(OrbxError):
        make_plan(1000, 1.2, 2, 20, 7, 100, 400)           # nIni == 0: the reference divides by zero


def test_adapter_compiles_against_reference_surface():
    """The drop-in C++ adapter keeps the reference class surface; compile-check it against the minicv shim."""
    import subprocess
    src = os.path.join(ROOT, "multimot_track_b200", "adapter")
    if not os.path.exists(os.path.join(src, "ORBextractor.cc")):
        pytest.skip("adapter not written yet")
    cmd = ["g++", "-std=c++11", "-fsyntax-only", "-I", os.path.join(ROOT, "oracle", "minicv"), "-I", os.path.join(ROOT, "oracle"),
           "-I", os.path.join(ROOT, "include"), "-I", src, os.path.join(src, "ORBextractor.cc"), os.path.join(src, "ORBmatcher_core.cc"),
           os.path.join(src, "adapter_check.cc")]
    subprocess.check_call(cmd)


def test_pool_entry_points_without_gpu(lib):
    """The dispatcher's host-only pieces: the frame split every submit uses equals multimot_track_b200.sharding (the split
    bench.py's ranks use), creation fails loudly without a device, NULL handling, struct sizes the C header promises."""
    import ctypes
    from multimot_track_b200 import _lib
    from multimot_track_b200.sharding import seam_pairs, shard_bounds
    for F in (1, 5, 13, 32, 64, 100000):
        for G in (1, 2, 3, 4, 8):
            covered = 0
            for g in range(G):
                first, count = ctypes.c_int(), ctypes.c_int()
                lib.orbx_pool_shard_range(F, G, g, ctypes.byref(first), ctypes.byref(count))
                assert (first.value, first.value + count.value) == shard_bounds(F, g, G)
                assert first.value == covered
                covered += count.value
            assert covered == F
            assert all(b == a + 1 for a, b in seam_pairs(F, G))
    assert ctypes.sizeof(_lib.OrbxShardResult) == 48 and ctypes.sizeof(_lib.OrbxPoolConfig) == 56 and ctypes.sizeof(_lib.OrbxConfig) == 36  # sizeof() of the C structs of include/orbx.h
    assert lib.orbx_pool_devices(None) == -1 and lib.orbx_pool_depth(None) == -1 and lib.orbx_pool_handle(None, 0, 0) is None
    assert lib.orbx_pool_collect(None, 0, None) == -1 and lib.orbx_pool_set_option(None, 1, 1) == -1
    assert lib.orbx_expand_keypoints(None, None, 0, None) == -1
    import torch
    if not torch.cuda.is_available():
        cfg = _lib.OrbxPoolConfig(_lib.OrbxConfig(1000, 1.2, 8, 20, 7, 0, 0, 0, -1), 0, None, 2, 0)
        p = ctypes.c_void_p()
        assert lib.orbx_pool_create(ctypes.byref(cfg), ctypes.byref(p)) == -3 and not p.value
        assert b"no usable CUDA device" in lib.orbx_pool_last_error(None)
    bad = _lib.OrbxPoolConfig(_lib.OrbxConfig(1000, 1.2, 8, 20, 7, 0, 0, 0, -1), -1, None, 2, 0)
    p = ctypes.c_void_p()
    assert lib.orbx_pool_create(ctypes.byref(bad), ctypes.byref(p)) == -1
