"""ctypes binding of libpgb200.so (the C-ABI declared in include/pgb200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C pilotguru_b200/csrc``.  There is no
fallback: if the shared object is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGB200_LIB") or os.path.join(_HERE, "libpgb200.so")   # PGB200_LIB: A/B builds of the same ABI

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])

PGB_OK, PGB_ERR_INVALID, PGB_ERR_CUDA, PGB_ERR_CAPACITY, PGB_ERR_NUMERIC = 0, -1, -2, -3, -4


class PgbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpgb200 error {code}: {msg}")
        self.code = code


_lib = None

vp = C.c_void_p
_SIGS = {
    "pgb_last_error": (C.c_char_p, []),
    "pgb_version": (C.c_int, []),
    "pgb_launch_count": (C.c_uint64, []),
    "pgb_orb_create": (vp, [C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "pgb_orb_destroy": (None, [vp]),
    "pgb_orb_levels": (C.c_int, [vp]),
    "pgb_orb_scale_factor": (C.c_float, [vp]),
    "pgb_orb_scale_factors": (C.c_int, [vp, vp, vp, vp, vp]),
    "pgb_orb_features_per_level": (C.c_int, [vp, vp]),
    "pgb_orb_max_keypoints": (C.c_int, [vp]),
    "pgb_orb_level_size": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp]),
    "pgb_orb_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, vp, vp, vp, C.c_int]),
    "pgb_orb_get_level": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
    "pgb_orb_get_score_map": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
    "pgb_orb_get_candidates": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp]),
    "pgb_orb_get_blurred_level": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp]),
    "pgb_orb_run_stage": (C.c_int, [vp, C.c_int]),
    "pgb_orb_stream": (vp, [vp]),
    "pgb_orb_check": (C.c_int, [vp]),
    "pgb_frames_to_gray": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                     C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_size_t, C.c_size_t, vp]),
    "pgb_frames_to_gray_rotated": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                             C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_size_t, C.c_size_t, vp]),
    "pgb_match_median_flow": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "pgb_event_create": (vp, [C.c_int]),
    "pgb_event_record": (C.c_int, [C.c_int, vp, vp]),
    "pgb_event_synchronize": (C.c_int, [C.c_int, vp]),
    "pgb_event_destroy": (None, [C.c_int, vp]),
    "pgb_video_open": (vp, [C.c_int, C.c_char_p]),
    "pgb_video_info": (C.c_int, [vp, vp, vp, vp, vp, vp]),
    "pgb_video_frame_span": (C.c_int, [vp, C.c_int64, vp, vp]),
    "pgb_video_read_rgb": (C.c_int, [vp, C.c_int64, C.c_int, vp, C.c_size_t, C.c_size_t, vp, vp]),
    "pgb_video_close": (None, [vp]),
    "pgb_synth_frames": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "pgb_device_count": (C.c_int, []),
    "pgb_device_malloc": (vp, [C.c_int, C.c_size_t]),
    "pgb_device_free": (None, [C.c_int, vp]),
    "pgb_host_malloc_pinned": (vp, [C.c_size_t]),
    "pgb_host_free_pinned": (None, [vp]),
    "pgb_memcpy_async": (C.c_int, [C.c_int, vp, vp, C.c_size_t, C.c_int, vp]),
    "pgb_stream_synchronize": (C.c_int, [C.c_int, vp]),
    "pgb_descriptor_distance": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp]),
    "pgb_matcher_create": (vp, [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, vp]),
    "pgb_matcher_destroy": (None, [vp]),
    "pgb_match_by_projection": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_float, vp, C.c_int, vp, vp, C.c_int]),
    "pgb_match_consecutive": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, vp,
                                        C.c_int, vp, vp]),
    "pgb_match_for_initialization": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_float,
                                               C.c_float, C.c_float, vp, vp, C.c_int]),
    "pgb_match_map_points": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_float, C.c_float,
                                       C.c_float, C.c_float, C.c_float, vp, C.c_int, vp, vp, C.c_int]),
    "pgb_match_by_bow": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                   vp, C.c_int, C.c_int, vp, vp, C.c_int]),
    "pgb_distinctive_descriptors": (C.c_int, [vp, vp, C.c_int, vp, C.c_int, vp]),
    "pgb_pose_optimization": (C.c_int, [C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_float,
                                        C.c_float, C.c_float, vp, vp, vp, C.c_int, vp]),
    "pgb_comm_unique_id": (C.c_int, [vp]),
    "pgb_comm_create": (vp, [C.c_int, C.c_int, C.c_int, vp]),
    "pgb_comm_create_all": (C.c_int, [C.c_int, vp, vp]),
    "pgb_comm_destroy": (None, [vp]),
    "pgb_comm_rank": (C.c_int, [vp]),
    "pgb_comm_size": (C.c_int, [vp]),
    "pgb_comm_nccl_version": (C.c_int, []),
    "pgb_allgather_feats": (C.c_int, [vp, vp, vp, C.c_size_t, vp]),
    "pgb_frame_record_bytes": (C.c_size_t, [C.c_int]),
    "pgb_frame_record_pack": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp]),
    "pgb_frame_record_unpack": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp]),
    "pgb_imu_create": (vp, [C.c_int, vp, vp, C.c_size_t, vp, vp, C.c_size_t, vp]),
    "pgb_imu_destroy": (None, [vp]),
    "pgb_imu_merged_count": (C.c_int64, [vp]),
    "pgb_imu_merged_events": (C.c_int, [vp, vp, vp, vp]),
    "pgb_imu_set_window": (C.c_int, [vp, vp, vp, C.c_int]),
    "pgb_imu_window_intervals": (C.c_int64, [vp]),
    "pgb_imu_eval": (C.c_int, [vp, vp, vp, vp]),
    "pgb_imu_minimize": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_double]),
    "pgb_imu_integrate": (C.c_int, [vp, vp, C.c_int64, vp, vp, vp, vp, vp, vp]),
    "pgb_imu_fit_windows": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                      vp, vp, vp, vp, vp]),
    "pgb_imu_num_windows": (C.c_int, [C.c_int, C.c_int]),
    "pgb_imu_last_kernel_ms": (C.c_int, [vp, vp, vp, vp, vp]),
    "pgb_imu_fit_windows_fwd": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                          vp, vp, vp, vp, vp, C.c_double, C.c_double, vp, vp]),
    "pgb_principal_rotation_axes": (C.c_int, [C.c_int, vp, vp, C.c_size_t, C.c_int64, vp, vp]),
    "pgb_angular_velocities_around_axis": (C.c_int, [C.c_int, vp, C.c_size_t, vp, vp]),
    "pgb_time_averaged_values": (C.c_int, [C.c_int, vp, vp, C.c_int64, vp, C.c_int64, vp, vp]),
    "pgb_smooth_time_series": (C.c_int, [C.c_int, vp, vp, C.c_int64, vp, C.c_int64, C.c_double, vp]),
}


def declared_symbols() -> list[str]:
    """Every function name include/pgb200.h declares (parsed from the header; used by the CPU ABI test)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_HERE), "include", "pgb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(", hdr)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PgbError(PGB_ERR_CUDA, f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise PgbError(rc, lib().pgb_last_error().decode("utf-8", "replace"))


def last_error() -> str:
    return lib().pgb_last_error().decode("utf-8", "replace")


def launch_count() -> int:
    return int(lib().pgb_launch_count())


def np_ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
