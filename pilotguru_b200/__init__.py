"""pilotguru_b200: B200-native (sm_100a) implementation of pilotguru's ORB extract + match + IMU calibration hot
path.  The compute lives in libpgb200.so (C-ABI in include/pgb200.h); this package is the thin host-side mirror of
the reference's call boundaries used by the tests and bench.py."""
from ._lib import KP_DTYPE, PgbError, launch_count  # noqa: F401
