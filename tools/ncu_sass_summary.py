#!/usr/bin/env python3
"""Summarise an ncu report's source page (SASS granularity) for one kernel: executed warp-instructions per region
(regions split at barriers / named marker opcodes) and the opcode mix.  Runs here (no GPU):
    python tools/ncu_sass_summary.py gpurun_out/x.ncu-rep [launch_index] [--list a b]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
launch = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].lstrip('-').isdigit() else -1
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
s = starts[launch]
end = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
hdr = rows[s + 1]
blk = [r for r in rows[s + 2:end] if len(r) > 5]
ie, ss, te = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Avg. Predicated-On Threads Executed')
tot = sum(int(r[ie]) for r in blk)
print(rows[s][1][:90], "total warp-inst", tot, "sass lines", len(blk))
if '--list' in sys.argv:
    k = sys.argv.index('--list')
    a, b = int(sys.argv[k + 1]), int(sys.argv[k + 2])
    for i in range(a, min(b, len(blk))):
        r = blk[i]
        print(i, int(r[ie]) // 1000, r[ss], r[te], r[1].strip()[:90])
    sys.exit(0)
marks = ('BAR.SYNC', 'SYNCS.PHASECHK', 'UTMALDG', 'UTMASTG', 'SHFL.IDX', 'EXIT', 'WARPSYNC')
acc = 0; start = 0; ops = collections.Counter()
for i, r in enumerate(blk):
    sass = r[1].strip()
    acc += int(r[ie])
    op = sass.split()[1] if sass.startswith('@') else sass.split()[0]
    ops[op.split('.')[0]] += int(r[ie])
    if any(m in sass for m in marks):
        print(f"  lines {start:4d}-{i:4d}  {acc/1e6:8.2f} M ({acc/tot:5.1%})  ends at {sass[:50]}")
        acc = 0; start = i + 1
print("opcode mix:", ", ".join(f"{k} {v/tot:.1%}" for k, v in ops.most_common(14)))

# ---- pipe classes per region (ALU pipe = LOP3/PRMT/SHF/VABSDIFF4/VIMNMX/VIADD/IADD3/ISETP/SEL/LEA...; FMA = IMAD*)
def pipe(op):
    b = op.split('.')[0]
    if b in ('IMAD', 'FFMA', 'FMUL', 'FADD', 'HFMA2', 'HMNMX2'): return 'fma'
    if b in ('LDS', 'STS', 'LDG', 'STG', 'LDC', 'LDCU', 'ATOMS', 'SHFL', 'LDSM', 'UTMALDG', 'UTMASTG', 'SYNCS', 'MEMBAR', 'FENCE'): return 'lsu'
    if b in ('BRA', 'BSSY', 'BSYNC', 'EXIT', 'WARPSYNC', 'NOP', 'BAR', 'VOTE', 'VOTEU', 'ELECT'): return 'ctl'
    if b in ('POPC', 'FLO', 'BREV', 'MUFU', 'I2F', 'F2I'): return 'xu'
    if b.startswith('U') or b in ('S2UR', 'R2UR', 'S2R', 'CS2R'): return 'uni'
    return 'alu'
acc = collections.Counter(); start = 0
print("region pipe mix (M warp-inst):")
for i, r in enumerate(blk):
    sass = r[1].strip()
    op = sass.split()[1] if sass.startswith('@') else sass.split()[0]
    acc[pipe(op)] += int(r[ie])
    if any(m in sass for m in marks):
        if sum(acc.values()) > 0.3e6:
            print(f"  lines {start:4d}-{i:4d} " + " ".join(f"{k}={v/1e6:.2f}" for k, v in sorted(acc.items())))
        acc = collections.Counter(); start = i + 1
