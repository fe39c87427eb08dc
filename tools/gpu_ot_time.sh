#!/bin/bash
# optical_trajectories on 10 000 device-rendered frames, one GPU: wall time of the frame loop under a few settings (run under gpurun)
cd "$(dirname "$0")/.."
python -c "
import sys; sys.path.insert(0,'.')
from pilotguru_b200 import synth
synth.canvas().tofile('/tmp/canvas.gray')
open('/tmp/settings.yml','w').write('%YAML:1.0\nCamera_fps: 30\nCamera_RGB: 1\nORBextractor_nFeatures: 1000\nORBextractor_scaleFactor: 1.2\nORBextractor_nLevels: 8\nORBextractor_iniThFAST: 20\nORBextractor_minThFAST: 7\n')
"
mkdir -p /tmp/ot_out
run() { echo "== $*"; env "$@" pilotguru_b200/host/optical_trajectories --vocabulary_file=unused --camera_settings /tmp/settings.yml --out_dir /tmp/ot_out --in_video=synth:/tmp/canvas.gray:2400x1400:10000:1920x1080 --batch=${BATCH:-64} --logtostderr 2>&1 | grep "extract+match"; }
run A=1
run A=1
run PGB_FC_NO_SIDE=1
BATCH=128 run A=1
BATCH=256 run A=1
