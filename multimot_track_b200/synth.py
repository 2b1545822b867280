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
