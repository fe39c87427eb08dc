#!/usr/bin/env python3
"""Debug aid: candidates of the fused FAST kernel vs the unfused pair, per level; prints the differing entries."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pilotguru_b200 import synth
from pilotguru_b200.orb import ORBextractor
w, h = 1920, 1080
frames = np.stack([synth.frame(t, w=w, h=h) for t in range(4)])
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=4)
ex.extract_batch(frames)
fused = [[ex.candidates(l, frame=f) for l in range(8)] for f in range(4)]
ex.run_stage(2); ex.check()
for f in range(4):
    for l in range(8):
        u = ex.candidates(l, frame=f)
        a = set(map(tuple, fused[f][l].tolist())); b = set(map(tuple, u.tolist()))
        if a != b or len(fused[f][l]) != len(u):
            print("frame", f, "level", l, "fused", len(fused[f][l]), "unfused", len(u), "only fused", sorted(a - b)[:10], "only unfused", sorted(b - a)[:10])
            lw, lh = ex.level_size(w, h, l)
            sm = ex.score_map(l, frame=f) if 'frame' in ex.score_map.__code__.co_varnames else None
            for (x, y, s) in sorted(a - b)[:3] + sorted(b - a)[:3]:
                print("  at", x, y, s, "level size", lw, lh)
                if sm is not None:
                    print(sm[y + 16 - 2:y + 16 + 3, x + 16 - 2:x + 16 + 3])
print("done")
