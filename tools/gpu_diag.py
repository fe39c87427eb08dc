#!/usr/bin/env python3
"""Stage-by-stage parity diagnostics of the CUDA ORB/match path against the oracle (runs on the GPU box).
Writes gpurun_out/diag.json.  Test infrastructure (uses oracle/)."""
import json, os, sys, time, traceback
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
from pilotguru_b200 import synth
from pilotguru_b200.orb import ORBextractor
from pilotguru_b200.matcher import ORBmatcher

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def case(name, img, nfeat=1000):
    r = {}
    h, w = img.shape
    ex = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=2)
    orc = O.OrbOracle(nfeat, 1.2, 8, 20, 7)
    ok, od = orc.extract(img)
    t = time.time(); gk, gd = ex(img); r["t_first_call_s"] = time.time() - t
    r["n_oracle"], r["n_gpu"] = len(ok), len(gk)
    for l in range(8):
        a = ex.image_pyramid(l); b = orc.level(l)
        r[f"pyr{l}"] = int((a != b).sum()) if a.shape == b.shape else f"shape {a.shape} vs {b.shape}"
        sm = ex.score_map(l); om = O.fast_score_map(b, 7)
        r[f"score{l}"] = int((sm != om).sum()) if sm.shape == om.shape else "shape"
        if r[f"score{l}"]:
            ys, xs = np.nonzero(sm != om)
            r[f"score{l}_ex"] = [(int(x), int(y), int(sm[y, x]), int(om[y, x])) for x, y in list(zip(xs, ys))[:5]]
        gc = ex.candidates(l); oc = orc.candidates(l)
        r[f"cand{l}"] = "equal" if np.array_equal(gc, oc) else f"gpu {len(gc)} oracle {len(oc)}"
        bl = ex.blurred_level(l); ob = O.gaussian_blur7(b)
        r[f"blur{l}"] = int((bl != ob).sum())
    if len(ok) == len(gk):
        for fld in KPF:
            r["kp_" + fld] = int((ok[fld] != gk[fld]).sum())
        r["desc_rows_diff"] = int((od != gd).any(axis=1).sum())
        r["desc_bits_diff"] = int(np.unpackbits(od ^ gd).sum())
    else:
        # compare per-level counts
        r["lvl_counts_gpu"] = np.bincount(gk["octave"], minlength=8).tolist()
        r["lvl_counts_oracle"] = np.bincount(ok["octave"], minlength=8).tolist()
    out[name] = r
    print(name, json.dumps(r), flush=True)
    ex.close()
    return ok, od


KPF = ["x", "y", "size", "angle", "response", "octave", "class_id"]
try:
    f0 = synth.frame(0)
    k0, d0 = case("synth1080_t0", f0)
    rng = np.random.default_rng(5)
    case("noise_640x480", rng.integers(0, 256, (480, 640), dtype=np.uint8))
    case("synth_odd_701x403", synth.frame(3, w=701, h=403))
    case("synth_small_nf300", synth.frame(1, w=400, h=300), nfeat=300)
    # batch + match
    frames = np.stack([synth.frame(t) for t in range(4)])
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)
    kps, desc, counts = ex.extract_batch(frames)
    orc = O.OrbOracle()
    r = {"counts": counts.tolist()}
    oks = []
    for t in range(4):
        ok, od = orc.extract(frames[t]); oks.append((ok, od))
        n = counts[t]
        r[f"f{t}_equal"] = bool(n == len(ok) and np.array_equal(kps[t, :n], ok) and np.array_equal(desc[t, :n], od))
    sf = ex.GetScaleFactors()
    m = ORBmatcher(0.9, True, max_feats=ex.cap, max_batch=4)
    for t in range(1, 4):
        (pk, pd), (ck, cd) = oks[t - 1], oks[t]
        fl = synth.flow(t)
        uv = np.stack([pk["x"] + np.float32(fl[0]), pk["y"] + np.float32(fl[1])], axis=1).astype(np.float32)
        for th in (15.0, 30.0):
            on, om, _ = O.search_by_projection(ck, cd, uv, pk["octave"], pk["angle"], pd, np.ones(len(pk), np.uint8),
                                               (0, 1920, 0, 1080), th, sf)
            gn, gm = m.SearchByProjection(ck, cd, uv, pk["octave"], pk["angle"], pd, np.ones(len(pk), np.uint8),
                                          (0.0, 1920.0, 0.0, 1080.0), th, sf)
            r[f"match_t{t}_th{int(th)}"] = dict(oracle=on, gpu=gn, equal=bool(np.array_equal(om, gm)))
    a = rng.integers(0, 256, (1000, 32), dtype=np.uint8); b = rng.integers(0, 256, (1000, 32), dtype=np.uint8)
    gdist = ORBmatcher.DescriptorDistance(a, b)
    odist = np.array([O.descriptor_distance(a[i], b[i]) for i in range(1000)])
    r["desc_distance_equal"] = bool(np.array_equal(gdist, odist))
    out["batch_match"] = r
    print("batch_match", json.dumps(r), flush=True)
    # timing of stages on a 32-frame batch
    import torch
    B = 32
    fr = np.stack([synth.frame(t) for t in range(B)])
    ex32 = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
    ex32.extract_batch(fr)
    st = torch.cuda.ExternalStream(ex32.stream)
    tm = {}
    with torch.cuda.stream(st):
        for which, nm in enumerate(["pyramid", "fast", "cells", "octree", "orient_desc"]):
            for _ in range(2): ex32.run_stage(which)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5): ex32.run_stage(which)
            e1.record(st); e1.synchronize()
            tm[nm + "_us_per_frame"] = e0.elapsed_time(e1) * 1000 / 5 / B
    out["stage_times_B32"] = tm
    print("stage_times", json.dumps(tm), flush=True)
except Exception:
    out["exception"] = traceback.format_exc()
    print(out["exception"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag.json"), "w"), indent=1)
