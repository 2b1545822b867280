#!/usr/bin/env python3
"""First-light diagnostics on the GPU box: stage-by-stage mismatch report of the CUDA
path against the oracle, written to gpurun_out/diag.txt (never raises on mismatches)."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s); out.write(s + "\n"); out.flush()


def main():
    import multimot_track_b200 as orb
    from multimot_track_b200.synth import value_noise_frame, uniform_noise_frame
    from oracle.oracle import Oracle
    cases = [("syn0", value_noise_frame(0, 375, 1242), (2000, 1.2, 8, 20, 7)),
             ("small", value_noise_frame(3, 240, 320), (500, 1.2, 4, 20, 7)),
             ("noise", uniform_noise_frame(5, 375, 1242), (2000, 1.2, 8, 20, 7))]
    for name, img, params in cases:
        P("==== case", name, img.shape, params)
        ext = orb.ORBextractor(*params)
        o = Oracle(*params)
        t = time.time(); kps, desc = ext(img); P("gpu extract %.1f ms (first call incl. allocation)" % ((time.time() - t) * 1e3))
        t = time.time(); kps, desc = ext(img); P("gpu extract %.2f ms (second call)" % ((time.time() - t) * 1e3))
        ok, od = o(img)
        for l in range(params[2]):
            gi, oi = ext.pyramid_level(l), o.level_image(l)
            pm = int((gi != oi).sum())
            gb, ob = ext.blurred_level(l), o.level_blurred(l)
            bm = int((gb != ob).sum()) if ob is not None else -1
            gc, oc = ext.candidates(l), o.level_candidates(l)
            sg, so = set(map(tuple, gc.tolist())), set(map(tuple, oc.tolist()))
            P("level %d size %s pyr_mismatch %d blur_mismatch %d cand gpu %d oracle %d only_gpu %d only_oracle %d order_equal %s nkp_oracle %d"
              % (l, ext.level_size(l), pm, bm, len(gc), len(oc), len(sg - so), len(so - sg), gc.shape == oc.shape and bool(np.array_equal(gc, oc)), o.level_nkeypoints(l)))
            if sg - so: P("   sample only_gpu", sorted(sg - so)[:5])
            if so - sg: P("   sample only_oracle", sorted(so - sg)[:5])
        P("keypoints gpu %d oracle %d per level gpu %s oracle %s" % (len(kps), len(ok), np.bincount(kps["octave"], minlength=params[2]).tolist() if len(kps) else [],
                                                                      np.bincount(ok["octave"], minlength=params[2]).tolist()))
        if len(kps) == len(ok):
            for f in ("x", "y", "size", "response", "octave", "class_id"):
                P("  field", f, "mismatches", int((kps[f] != ok[f]).sum()))
            d = np.abs(kps["angle"].astype(np.float64) - ok["angle"]); d = np.minimum(d, 360 - d)
            P("  angle max abs diff deg %.3g, bit-identical %d / %d" % (d.max(), int((kps["angle"].view(np.uint32) == ok["angle"].view(np.uint32)).sum()), len(kps)))
            P("  descriptor bits differing %d of %d; rows differing %d" % (int(np.unpackbits(desc ^ od).sum()), desc.size * 8, int((desc != od).any(axis=1).sum())))
        else:
            sg = set(zip(kps["x"].tolist(), kps["y"].tolist(), kps["octave"].tolist())); so = set(zip(ok["x"].tolist(), ok["y"].tolist(), ok["octave"].tolist()))
            P("  set sym diff", len(sg ^ so))
        m = orb.ORBmatcher(0.9, extractor=ext)
        B = od[::-1].copy()
        idx, d1, d2, acc = m.match(od, B, 100, 0.9)
        oi_, o1, o2, oa = Oracle.match(od, B, 100, 0.9)
        P("matcher mismatches idx %d d1 %d d2 %d acc %d (accepted %d)" % ((idx != oi_).sum(), (d1 != o1).sum(), (d2 != o2).sum(), (acc != oa).sum(), acc.sum()))
        P("launches", ext.launch_count)


try:
    main()
except Exception:
    P(traceback.format_exc())
out.close()
