python - <<'PY'
import sys; sys.path.insert(0,'.')
from pilotguru_b200 import synth
synth.canvas().tofile('/tmp/canvas.gray')
open('/tmp/s.yml','w').write("%YAML:1.0\nCamera_fps: 30\nORBextractor_nFeatures: 1000\n")
PY
nproc
for b in ot_inline_ab optical_trajectories; do for i in 1 2; do mkdir -p /tmp/o_$b; ./pilotguru_b200/host/$b --vocabulary_file=x --camera_settings /tmp/s.yml --out_dir /tmp/o_$b --in_video=synth:/tmp/canvas.gray:2400x1400:10000:1920x1080 --batch=128 --logtostderr 2>&1 | grep "extract+match"; done; done
cmp /tmp/o_ot_inline_ab/trajectory-0.json /tmp/o_optical_trajectories/trajectory-0.json && echo same
