"""Synthetic frames of SURVEY.md section 8d: the value-noise generator whose FAST
corner density is close to the KITTI sample's.  Needs cv2 (in the image) only for the
INTER_CUBIC up-sampling of the coarse grid; CRC32 anchors in tests/golden pin it."""
import numpy as np


def value_noise_frame(seed, height, width):
    import cv2
    rng = np.random.default_rng(seed)
    rows, cols = -(-height // 12) + 2, -(-width // 12) + 2
    coarse = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
    up = cv2.resize(coarse, (12 * cols, 12 * rows), interpolation=cv2.INTER_CUBIC)[12:12 + height, 12:12 + width]
    noise = rng.integers(-4, 5, (height, width), dtype=np.int16)
    return np.clip(up.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def frame_pool(height, width, count=64, first_seed=0):
    return np.stack([value_noise_frame(first_seed + i, height, width) for i in range(count)])


def uniform_noise_frame(seed, height, width):
    return np.random.default_rng(seed).integers(0, 256, (height, width), dtype=np.uint8)


def stereo_pair(seed, h, w, max_disp=60.0):
    """A rectified stereo pair for Frame::ComputeStereoMatches tests: the right image is the left one resampled
    with a smooth positive disparity field d(x, y) in [2, max_disp] (right(x) = left(x + d)), plus +-2 of fresh noise.
    The images are 16 px wider internally so the shifted samples exist."""
    import cv2
    left_wide = value_noise_frame(seed, h, w + int(max_disp) + 8)
    rng = np.random.default_rng(seed + 7919)
    coarse = rng.random((h // 64 + 3, w // 64 + 3)).astype(np.float32)
    disp = cv2.resize(coarse, (w, h), interpolation=cv2.INTER_CUBIC)
    disp = 2.0 + (max_disp - 2.0) * np.clip(disp, 0, 1)
    xs = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0)
    ys = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1)
    right = cv2.remap(left_wide, xs + disp, ys, cv2.INTER_LINEAR)
    right = np.clip(right.astype(np.int16) + rng.integers(-2, 3, (h, w), dtype=np.int16), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left_wide[:, :w]), right
