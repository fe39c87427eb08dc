import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pilotguru_b200 import synth
from pilotguru_b200.orb import ORBextractor
B = int(os.environ.get("CHK_BATCH", 16))
frames = np.stack([synth.frame(t) for t in range(B)])
ex = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=B)
ex.extract_batch(frames)
for _ in range(3):
    ex.run_stage(1)
ex.check()
print("done")
