#!/usr/bin/env python3
"""Single-frame latency of the extractor (BASELINE configs[1]): host call vs device-resident call vs per-stage, wall clock."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pilotguru_b200 import synth
from pilotguru_b200.orb import ORBextractor

f = synth.frame(0)
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
def med(fn, n=40):
    t = []
    for i in range(n + 5):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    return 1e6 * float(np.median(t[5:]))
print("host call (pageable in, host out): %.0f us" % med(lambda: ex(f)))
pin = torch.from_numpy(f).pin_memory().numpy()
print("host call (pinned in): %.0f us" % med(lambda: ex(pin)))
d = torch.from_numpy(f).cuda()
cap = ex.cap
kps = torch.zeros((1, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((1, cap, 32), dtype=torch.uint8, device="cuda")
cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
def resident():
    ex.extract_ptr(d.data_ptr(), 3, 1, 1920, 1080, 1920, 1920 * 1080, kps.data_ptr(), desc.data_ptr(), cnt.data_ptr(), cap)
    ex.check()
print("device-resident call + sync: %.0f us" % med(resident))
for name, s in (("pyramid", 0), ("fast_cells", 1), ("octree", 3), ("orient_desc", 4)):
    def st():
        ex.run_stage(s); ex.check()
    print("  stage %-12s + sync: %.0f us" % (name, med(st)))
