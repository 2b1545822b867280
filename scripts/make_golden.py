#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ (run in the build
container, where /root/reference and cv2 4.13.0 exist; the GPU box has neither the
reference tree nor needs cv2 for these).

  kitti_gray_00000{0..4}.png   grey KITTI sample frames, converted exactly like the
                              reference driver (imread UNCHANGED + RGB2GRAY on the stored
                              BGR bytes: Examples/RGB-D/rgbd_tum.cc:122, src/Tracking.cc:459-465)
  golden_kitti.npz            outputs of the REFERENCE's own ORBextractor.cc (canonical
                              tie-break build, oracle/_ref/liborbref_canon.so) on those frames
                              for kitti03.yaml's parameters (4000,1.2,8,20,7) and the
                              benchmark's (2000,...): keypoints, descriptors, pyramid CRC32s,
                              plus as-is (pointer tie-break) counts for the record
  golden_prims.npz            cv2 4.13.0 outputs for the OpenCV primitives the path uses
                              (resize, GaussianBlur, FAST, fastAtan2, copyMakeBorder)
  golden_synth.npz            CRC32 anchors of the synthetic value-noise frames + reference
                              keypoint counts on them
"""
import os
import sys
import zlib

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle, RefExtractor, build  # noqa: E402
from multimot_track_b200.synth import value_noise_frame, uniform_noise_frame  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
build(ref=True)


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


# ---- KITTI frames through the reference's compiled extractor
g = {}
grays = []
for i in range(5):
    img = cv2.imread("%s/kitti_sample/image/%06d.png" % (REF, i), cv2.IMREAD_UNCHANGED)
    gray = cv2.cvtColor(img, cv2.COLOR_RGB2GRAY)
    cv2.imwrite(os.path.join(OUT, "kitti_gray_%06d.png" % i), gray, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    grays.append(gray)
    g["gray_crc_%d" % i] = np.uint32(crc(gray))
    for nfeat in (2000, 4000):
        ref = RefExtractor(nfeat, 1.2, 8, 20, 7, "canon")
        kps, desc = ref(gray)
        tag = "f%d_n%d" % (i, nfeat)
        g["kps_" + tag] = kps
        g["desc_" + tag] = desc
        g["pyr_crc_" + tag] = np.array([crc(ref.pyramid_level(l)) for l in range(8)], np.uint32)
        g["pyr_border_crc_" + tag] = np.array([crc(ref.pyramid_level(l, True)) for l in range(8)], np.uint32)
        asis = RefExtractor(nfeat, 1.2, 8, 20, 7, "asis")
        ka, _ = asis(gray)
        sa = set(zip(kps["x"].tolist(), kps["y"].tolist(), kps["octave"].tolist()))
        sb = set(zip(ka["x"].tolist(), ka["y"].tolist(), ka["octave"].tolist()))
        g["asis_n_" + tag] = np.int32(len(ka))
        g["asis_symdiff_" + tag] = np.int32(len(sa ^ sb))
        # stage anchors from the port (identical to the reference on these frames, see tests)
        o = Oracle(nfeat, 1.2, 8, 20, 7)
        o(gray)
        g["ncand_" + tag] = np.array([len(o.level_candidates(l)) for l in range(8)], np.int32)
        g["mincells_" + tag] = np.array([o.level_min_cells(l)[0] for l in range(8)], np.int32)
        g["blur_crc_" + tag] = np.array([crc(o.level_blurred(l)) for l in range(8)], np.uint32)
# Config 1's consecutive-frame matching: frame i against frame i+1 with the reference's scan (src/ORBmatcher.cc:574-605),
# distances by the reference's own compiled DescriptorDistance on a sample of pairs, TH_LOW / TH_HIGH, ratio 0.9
for i in range(4):
    dA, dB = g["desc_f%d_n4000" % i], g["desc_f%d_n4000" % (i + 1)]
    for th in (50, 100):
        idx, d1, d2, acc = Oracle.match(dA, dB, th, 0.9)
        for j in range(0, len(dA), 97):                    # spot-check the port's distances against the reference's function
            assert RefExtractor.descriptor_distance(dA[j], dB[idx[j]]) == d1[j]
        g["match_f%d_th%d" % (i, th)] = np.stack([idx, d1, d2, acc.astype(np.int32)], 1).astype(np.int16)
t = RefExtractor(2000, 1.2, 8, 20, 7).tables()
for k, v in t.items():
    g["tables2000_" + k] = v
t = RefExtractor(10000, 1.2, 12, 20, 7).tables()
for k, v in t.items():
    g["tables10000x12_" + k] = v
np.savez_compressed(os.path.join(OUT, "golden_kitti.npz"), **g)

# ---- OpenCV primitives (cv2 4.13.0)
p = {"cv2_version": np.array(cv2.__version__)}
rng = np.random.default_rng(123)
src = value_noise_frame(3, 97, 133)
p["resize_src"] = src
for k, (dw, dh) in enumerate([(111, 81), (92, 67), (133, 97), (64, 64), (200, 150)]):
    p["resize_dst_%d" % k] = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR)
p["blur_src"] = rng.integers(0, 256, (41, 67), dtype=np.uint8)
p["blur_dst"] = cv2.GaussianBlur(p["blur_src"], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
p["blur_small_src"] = rng.integers(0, 256, (5, 9), dtype=np.uint8)
p["blur_small_dst"] = cv2.GaussianBlur(p["blur_small_src"], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
p["border_dst"] = cv2.copyMakeBorder(p["blur_src"], 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
fast_src = np.concatenate([value_noise_frame(5, 60, 90), uniform_noise_frame(6, 60, 90)], axis=0)
p["fast_src"] = fast_src
for th in (20, 7):
    for nms in (1, 0):
        det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=bool(nms), type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        kp = det.detect(fast_src)
        p["fast_t%d_nms%d" % (th, nms)] = np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in kp], np.int32).reshape(-1, 3)
ys = rng.integers(-2 ** 22, 2 ** 22, 4000).astype(np.float32)
xs = rng.integers(-2 ** 22, 2 ** 22, 4000).astype(np.float32)
ys[:8] = 0
xs[4:12] = 0
p["atan2_y"], p["atan2_x"] = ys, xs
p["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in zip(ys, xs)], np.float32)
np.savez_compressed(os.path.join(OUT, "golden_prims.npz"), **p)

# ---- synthetic anchors
s = {}
for (h, w) in ((375, 1242), (1080, 1920), (2160, 3840)):
    s["crc_%dx%d_seed0" % (w, h)] = np.uint32(crc(value_noise_frame(0, h, w)))
for seed in (0, 1):
    fr = value_noise_frame(seed, 375, 1242)
    ref = RefExtractor(2000, 1.2, 8, 20, 7, "canon")
    kps, desc = ref(fr)
    s["kps_1242x375_seed%d" % seed] = kps
    s["desc_1242x375_seed%d" % seed] = desc
np.savez_compressed(os.path.join(OUT, "golden_synth.npz"), **s)
for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))
