#!/usr/bin/env python3
"""Breaks the e2e step of bench.py into legs (CUDA events + host wall clock) to see where the time beyond the bare
H2D transfer goes.  PGB_H2D_CHUNK is honoured by the library."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pilotguru_b200 import synth
from pilotguru_b200.matcher import ORBmatcher
from pilotguru_b200.orb import ORBextractor

B = int(os.environ.get("PGB_PROFILE_BATCH", 64)); W, H = 1920, 1080
host = torch.from_numpy(np.stack([synth.frame(t) for t in range(B)])).pin_memory()
dev = host.cuda()
flows = torch.from_numpy(np.array([synth.flow(t) for t in range(B)], np.float32)).cuda()
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
cap = ex.cap
st = torch.cuda.ExternalStream(ex.stream)
kps = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda"); desc = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
counts = torch.zeros(B, dtype=torch.int32, device="cuda"); match = torch.full((B, cap), -1, dtype=torch.int32, device="cuda")
nmatch = torch.zeros(B, dtype=torch.int32, device="cuda")
hk = torch.zeros((B, cap, 7), dtype=torch.float32).pin_memory(); hd = torch.zeros((B, cap, 32), dtype=torch.uint8).pin_memory()
hm = torch.zeros((B, cap), dtype=torch.int32).pin_memory(); hc = torch.zeros(B, dtype=torch.int32).pin_memory()
mt = ORBmatcher(0.9, True, max_feats=cap, max_batch=B, stream=ex.stream)
sf = ex.GetScaleFactors()

def extract(where, ptr): ex.extract_ptr(ptr, where, B, W, H, W, W * H, kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), cap)
def domatch(): mt.match_consecutive_ptr(B - 1, cap, kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), flows[1:].contiguous().data_ptr(), float(W), float(H), 15.0, sf, match.data_ptr(), nmatch.data_ptr())
def d2h():
    hk.copy_(kps, non_blocking=True); hd.copy_(desc, non_blocking=True); hm.copy_(match, non_blocking=True); hc.copy_(counts, non_blocking=True)

def timed(name, fn, n=10):
    with torch.cuda.stream(st):
        for _ in range(2): fn(); st.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        host_enq = 0.0
        torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record(st)
        for _ in range(n):
            t1 = time.perf_counter(); fn(); host_enq += time.perf_counter() - t1
            st.synchronize()
        e1.record(st); e1.synchronize(); wall = time.perf_counter() - t0
    print(f"{name:44s} events {e0.elapsed_time(e1)/n:7.3f} ms  wall {wall/n*1e3:7.3f} ms  host enqueue {host_enq/n*1e3:7.3f} ms", flush=True)

with torch.cuda.stream(st):
    timed("H2D only (one 133 MB copy)", lambda: dev.copy_(host, non_blocking=True))
    timed("H2D only (64 x 2 MB copies)", lambda: [dev[i].copy_(host[i], non_blocking=True) for i in range(B)])
    timed("extract, frames resident", lambda: extract(3, dev.data_ptr()))
    timed("extract from pinned host (pipelined)", lambda: extract(2, host.data_ptr()))
    timed("... + match", lambda: (extract(2, host.data_ptr()), domatch()))
    timed("... + match + D2H", lambda: (extract(2, host.data_ptr()), domatch(), d2h()))
    timed("match only", domatch)
    timed("D2H only", d2h)
