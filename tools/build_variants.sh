#!/bin/bash
# A/B builds of libpgb200.so with different tile / occupancy constants (-DPGB_*), into pilotguru_b200/variants/.
# Run one with PGB200_LIB=pilotguru_b200/variants/<name>.so python bench.py --no-cpu-baseline --no-calibration
set -e
cd "$(dirname "$0")/../pilotguru_b200/csrc"
mkdir -p ../variants
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -cudart static --expt-relaxed-constexpr"
build() { name=$1; shift; /usr/local/cuda/bin/nvcc $FLAGS $@ -shared -o ../variants/$name.so *.cu -ldl -lpthread -lrt & }
if [ $# -gt 0 ]; then   # build_variants.sh name -DFLAG=... [name2 -DFLAG...]: one variant per (name, flag) pair
  while [ $# -gt 1 ]; do build "$1" "$2"; shift 2; done
else
  build cw2 -DPGB_CELL_WARPS=2
  build cw8 -DPGB_CELL_WARPS=8
  build py16 -DPGB_PY_H=16
  build occ6 -DPGB_FS_OCC=6
fi
wait
ls -la ../variants
