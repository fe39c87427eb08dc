#!/bin/bash
# Where the wall time of the C5 host binaries goes (1 GPU, 30-minute inputs): run under gpurun.
set -e
cd "$(dirname "$0")/.."
W=/tmp/pgb_c5t; mkdir -p $W
python - <<'P'
import sys, time, json
sys.path.insert(0, '.')
from pilotguru_b200 import synth
d = synth.imu_gps(1800.0, 500.0)
synth.write_imu_gps_json(d, '/tmp/pgb_c5t')
json.dump({"frames": [{"frame_id": i, "time_usec": int(round(i * 1e6 / 30.0))} for i in range(54000)]}, open('/tmp/pgb_c5t/frames.json', 'w'))
P
H=pilotguru_b200/host
export PGB_IMU_TIMING=1
time $H/fit_motion --rotations_json $W/rotations.json --accelerations_json $W/accelerations.json --locations_json $W/locations.json --velocities_out_json $W/vel.json --steering_out_json $W/steer.json --forward_axis_out_json $W/fwd.json --logtostderr 2>&1 | grep -v "Sliding window" | tail -25
time $H/annotate_frames --frames_json $W/frames.json --in_json $W/vel.json --json_root_element_name velocities --json_value_name speed_m_s --out_json $W/vel_frames.json 2>&1 | tail -5
ls -la $W
