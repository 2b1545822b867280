"""CPU checks of the integer identities k_fast_fused (multimot_track_b200/csrc/kernels.cu, ff_process) is built on, against the
definition of OpenCV's cornerScore<16> that oracle/cvprim.c restates: the 32-operation min / max network per polarity, the
centre / threshold fold into one 32-bit multiply-add on packed halfwords, and the packed survivor entries.  The kernel itself is
compared with the oracle on the GPU (tests/test_gpu_parity.py); these run everywhere and pin the algebra."""
import numpy as np
import pytest


def _rings(n, seed):
    rng = np.random.default_rng(seed)
    e = rng.integers(0, 256, size=(n, 16)).astype(np.int32)
    e[: n // 4] = np.sort(e[: n // 4], axis=1)                      # monotone rings: long arcs
    e[n // 4: n // 2] = (e[n // 4: n // 2] > 127) * 255              # two-valued rings: many ties
    e[n // 2: n // 2 + 16] = np.eye(16, dtype=np.int32) * 200        # one bright pixel
    return e


def _brute(e):
    arcs = [np.stack([e[:, (k + t) % 16] for t in range(9)], 0) for k in range(16)]
    return np.max(np.stack([a.min(0) for a in arcs], 0), 0), np.min(np.stack([a.max(0) for a in arcs], 0), 0)


def _network(e):
    """The kernel's decomposition, operation for operation (three-input min / max written as nested two-input ones)."""
    E = lambda i: e[:, i % 16]
    lo2 = [np.minimum(E(2 * i + 1), E(2 * i + 2)) for i in range(8)]
    hi2 = [np.maximum(E(2 * i + 1), E(2 * i + 2)) for i in range(8)]
    mx = [np.maximum(E(2 * i), E(2 * i + 9)) for i in range(8)]
    mn = [np.minimum(E(2 * i), E(2 * i + 9)) for i in range(8)]
    bv, dv, ops = [None] * 8, [None] * 8, 32                          # 8 lo2 + 8 mx + 4 cmin + 8 bv + 4 for the final max, per polarity
    for i in range(4):
        cmin = np.minimum(np.minimum(lo2[2 * i + 1], lo2[(2 * i + 2) % 8]), lo2[(2 * i + 3) % 8])
        cmax = np.maximum(np.maximum(hi2[2 * i + 1], hi2[(2 * i + 2) % 8]), hi2[(2 * i + 3) % 8])
        bv[2 * i] = np.minimum(np.minimum(cmin, lo2[2 * i]), mx[2 * i])
        bv[2 * i + 1] = np.minimum(np.minimum(cmin, lo2[(2 * i + 4) % 8]), mx[2 * i + 1])
        dv[2 * i] = np.maximum(np.maximum(cmax, hi2[2 * i]), mn[2 * i])
        dv[2 * i + 1] = np.maximum(np.maximum(cmax, hi2[(2 * i + 4) % 8]), mn[2 * i + 1])
    return np.max(np.stack(bv), 0), np.min(np.stack(dv), 0), ops


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_arc_network_equals_the_sixteen_arcs(seed):
    e = _rings(40000, seed)
    maxmin, minmax = _brute(e)
    b, d, ops = _network(e)
    assert ops == 32 and (b == maxmin).all() and (d == minmax).all()


def test_network_gives_the_opencv_corner_score():
    """score = max(max_k min(arc_k) - c, c - min_k max(arc_k)) - 1, clamped at 0 -- cornerScore<16> evaluated on differences."""
    rng = np.random.default_rng(7)
    e = _rings(20000, 3)
    c = rng.integers(0, 256, size=e.shape[0]).astype(np.int32)
    b, d, _ = _network(e)
    score = np.maximum(np.maximum(b - c, c - d) - 1, 0)
    # OpenCV's form: the largest t such that 9 contiguous ring pixels are all > c + t or all < c - t  <=>  t <= score
    diff = e - c[:, None]
    ref = np.zeros_like(c)
    for k in range(16):
        arc = np.stack([diff[:, (k + t) % 16] for t in range(9)], 0)
        ref = np.maximum(ref, np.maximum(arc.min(0), -arc.max(0)) - 1)
    assert (score == np.maximum(ref, 0)).all()


def test_threshold_fold_has_no_borrow_between_halves():
    """(-x - th) on both halfwords of a packed register by ONE 32-bit multiply-add x * 0xffffffff + kv, for 8-bit x and th in 1..255."""
    x = np.arange(256, dtype=np.uint64)
    lo, hi = np.meshgrid(x, x)
    packed = (hi << np.uint64(16) | lo).ravel()
    for th in (1, 2, 7, 20, 128, 255):
        kv = np.uint64((0xFFFFFFFF - 0x00010001 * (th - 1)) & 0xFFFFFFFF)
        r = (packed * np.uint64(0xFFFFFFFF) + kv) & np.uint64(0xFFFFFFFF)
        want_lo = (-(packed & np.uint64(0xFFFF)).astype(np.int64) - th) & 0xFFFF
        want_hi = (-(packed >> np.uint64(16)).astype(np.int64) - th) & 0xFFFF
        assert ((r & np.uint64(0xFFFF)).astype(np.int64) == want_lo).all() and ((r >> np.uint64(16)).astype(np.int64) == want_hi).all()


def test_survivor_entry_round_trip():
    """entry = S'0 | row << 9 | S'1 << 16 (S' <= 255, row <= 63): the epilogue's decode and its packed counters."""
    rng = np.random.default_rng(5)
    s0, s1 = rng.integers(0, 256, 5000), rng.integers(0, 256, 5000)
    row = rng.integers(0, 64, 5000)
    en = (s0 | (s1 << 16)) + (row << 9)
    assert ((en & 0x1FF) == s0).all() and ((en >> 16) == s1).all() and (((en & 0xFE00) << 3) == (row << 12)).all()
    for s_ini in (1, 14, 40):                                        # S' >= s_ini  <=>  score >= iniThFAST
        e2 = en & 0x01FF01FF
        halves = np.stack([e2 & 0xFFFF, e2 >> 16], 0).astype(np.int64)
        strong = np.maximum(np.minimum(halves + (1 - s_ini), 1), 0)  # VIADDMNMX.S16x2.RELU(e2, 1 - s_ini, 1)
        pos = np.minimum(halves, 1)                                  # VIMNMX.U16x2(e2, 1)
        assert (strong[0] == (s0 >= s_ini)).all() and (strong[1] == (s1 >= s_ini)).all()
        assert (pos[0] == (s0 > 0)).all() and (pos[1] == (s1 > 0)).all()
