#!/usr/bin/env python3
"""A/B of the FAST score kernels (PGB_FAST_IMPL=v1|v2): parity of score maps / keypoints against the oracle and
per-launch timing of every stage on a 64-frame 1080p batch.  Test infrastructure (uses oracle/)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import torch
from pilotguru_b200 import synth
from pilotguru_b200.orb import ORBextractor

B = int(os.environ.get("CHK_BATCH", 64))
res = {}
rng = np.random.default_rng(5)
cases = {"synth1080": (synth.frame(0), 1000), "noise640": (rng.integers(0, 256, (480, 640), dtype=np.uint8), 1000),
         "odd701x403": (synth.frame(3, w=701, h=403), 777), "tiny300x200": (synth.frame(4, w=300, h=200), 300),
         "wide2000x300": (synth.frame(1, w=2000, h=300, cw=2480, ch=620), 500)}
refs = {}
for name, (img, nf) in cases.items():
    orc = O.OrbOracle(nf, 1.2, 8, 20, 7)
    ok, od = orc.extract(img)
    refs[name] = (ok, od, [O.fast_score_map(orc.level(l), 7) for l in range(8)])
frames = np.stack([synth.frame(t) for t in range(B)])
for impl in os.environ.get("CHK_IMPLS", "v1,v2").split(","):
    os.environ["PGB_FAST_IMPL"] = impl
    r = {}
    for name, (img, nf) in cases.items():
        h, w = img.shape
        ex = ORBextractor(nf, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
        gk, gd = ex(img)
        ok, od, maps = refs[name]
        bad = [int((ex.score_map(l) != maps[l]).sum()) for l in range(8)]
        r[name] = dict(score_mismatch=bad, kp_equal=bool(len(gk) == len(ok) and np.array_equal(gk, ok) and np.array_equal(gd, od)))
        if any(bad):
            l = [i for i, b in enumerate(bad) if b][0]
            sm = ex.score_map(l); ys, xs = np.nonzero(sm != maps[l])
            r[name]["examples"] = [(l, int(x), int(y), int(sm[y, x]), int(maps[l][y, x])) for x, y in list(zip(xs, ys))[:8]]
        ex.close()
    ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
    k, d, c = ex.extract_batch(frames)
    ok, od, _ = refs["synth1080"]
    r["batch_frame0_equal"] = bool(c[0] == len(ok) and np.array_equal(k[0, :c[0]], ok) and np.array_equal(d[0, :c[0]], od))
    st = torch.cuda.ExternalStream(ex.stream)
    tm = {}
    with torch.cuda.stream(st):
        for which, nm in enumerate(["pyramid", "fast", "cells", "octree", "orient_desc"]):
            for _ in range(3): ex.run_stage(which)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10): ex.run_stage(which)
            e1.record(st); e1.synchronize()
            tm[nm] = round(e0.elapsed_time(e1) * 1000 / 10 / B, 3)
    r["us_per_frame"] = tm
    r["fast_GBps"] = round(2 * 6419321 / (tm["fast"] * 1e-6) / 1e9, 1)
    ex.close()
    res[impl] = r
    print(impl, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "fast_check.json"), "w"), indent=1)
