#!/usr/bin/env python3
"""One pass of every stage kernel (pyramid x7, FAST, cell NMS, octree, orientation+descriptor, match) over a resident
batch, bracketed by cudaProfilerStart/Stop so that
    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/stages python tools/gpu_stage_profile.py
captures exactly those launches at the batch size the bench uses.  PGB_PROFILE_BATCH (default 32) frames."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pilotguru_b200 import synth
from pilotguru_b200.matcher import ORBmatcher
from pilotguru_b200.orb import ORBextractor

B = int(os.environ.get("PGB_PROFILE_BATCH", 32))
W, H = 1920, 1080
frames = torch.from_numpy(np.stack([synth.frame(t) for t in range(B)])).cuda()
flows = torch.from_numpy(np.array([synth.flow(t) for t in range(B)], np.float32)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
cap = ex.cap
kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
counts = torch.zeros(B, dtype=torch.int32, device="cuda")
match = torch.full((B, cap), -1, dtype=torch.int32, device="cuda")
nmatch = torch.zeros(B, dtype=torch.int32, device="cuda")
mt = ORBmatcher(0.9, True, max_feats=cap, max_batch=B, stream=ex.stream)
sf = ex.GetScaleFactors()
torch.cuda.synchronize()
ex.extract_ptr(frames.data_ptr(), 3, B, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), cap)
ex.check()
def do_match():
    mt.match_consecutive_ptr(B - 1, cap, kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), flows[1:].contiguous().data_ptr(),
                             float(W), float(H), 15.0, sf, match.data_ptr(), nmatch.data_ptr())
do_match(); ex.check()
torch.cuda.profiler.start()
for which in range(5):
    ex.run_stage(which)
do_match()
ex.check()
torch.cuda.profiler.stop()
print("profiled stages over", B, "frames; keypoints/frame", float(counts.float().mean()), "matches/pair", float(nmatch[:B - 1].float().mean()))
