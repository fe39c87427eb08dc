"""Host-side mirror of ORB_SLAM2::ORBextractor (thirdparty/orb-slam2/include/ORBextractor.h:51-85) over the
libpgb200 C-ABI.  Same constructor arguments, getters and call operator; adds the batched / device-resident
entry points the B200 path is built around."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, lib, np_ptr


class ORBextractor:
    """``ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)`` (ORBextractor.cc:410-470).

    ``extractor(image)`` returns ``(keypoints, descriptors)`` like ``operator()(image, mask, keypoints,
    descriptors)`` (ORBextractor.cc:1042-1104): keypoints as a structured array with cv::KeyPoint's fields,
    descriptors as an (N, 32) uint8 array.  The mask argument is ignored, as in the reference.
    """

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int,
                 max_width: int = 1920, max_height: int = 1080, max_batch: int = 1, device: int = 0, stream=None):
        self._h = None
        h = lib().pgb_orb_create(device, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height,
                                 max_batch, stream)
        if not h:
            raise _lib.PgbError(-1, _lib.last_error())
        self._h = C.c_void_p(h)
        self.nfeatures, self.scaleFactor, self.nlevels = nfeatures, scaleFactor, nlevels
        self.iniThFAST, self.minThFAST = iniThFAST, minThFAST
        self.max_batch = max_batch
        self.device = device
        self.cap = lib().pgb_orb_max_keypoints(self._h)
        L = nlevels
        self._scale = np.empty(L, np.float32); self._inv = np.empty(L, np.float32)
        self._s2 = np.empty(L, np.float32); self._is2 = np.empty(L, np.float32)
        check(lib().pgb_orb_scale_factors(self._h, np_ptr(self._scale), np_ptr(self._inv), np_ptr(self._s2),
                                          np_ptr(self._is2)))
        self._nper = np.empty(L, np.int32)
        check(lib().pgb_orb_features_per_level(self._h, np_ptr(self._nper)))

    def close(self):
        if self._h:
            lib().pgb_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters (ORBextractor.h:63-85)
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self._scale.copy()
    def GetInverseScaleFactors(self): return self._inv.copy()
    def GetScaleSigmaSquares(self): return self._s2.copy()
    def GetInverseScaleSigmaSquares(self): return self._is2.copy()
    def features_per_level(self): return self._nper.copy()

    @property
    def stream(self) -> int:
        return int(lib().pgb_orb_stream(self._h) or 0)

    # ---- operator()
    def __call__(self, image: np.ndarray, mask=None):
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("image must be CV_8UC1 (2-D uint8)")  # assert at ORBextractor.cc:1049
        kps, desc, counts = self.extract_batch(image[None])
        n = int(counts[0])
        return kps[0, :n].copy(), desc[0, :n].copy()

    def extract_batch(self, frames: np.ndarray):
        """frames: (n, H, W) uint8 host array.  Returns (kps[n, cap], desc[n, cap, 32], counts[n]) host arrays."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n, h, w = frames.shape
        kps = np.zeros((n, self.cap), KP_DTYPE)
        desc = np.zeros((n, self.cap, 32), np.uint8)
        counts = np.zeros(n, np.int32)
        check(lib().pgb_orb_extract(self._h, np_ptr(frames), 0, n, w, h, w, w * h, np_ptr(kps), np_ptr(desc),
                                    np_ptr(counts), self.cap))
        return kps, desc, counts

    IN_DEVICE, OUT_DEVICE = 1, 2

    def extract_ptr(self, gray_ptr: int, where: int, n: int, w: int, h: int, pitch: int, frame_stride: int,
                    kps_ptr: int, desc_ptr: int, counts_ptr: int, cap: int):
        """Raw-pointer form (device or pinned-host buffers owned by the caller, e.g. torch tensors).
        where: 0 host->host, OUT_DEVICE host frames -> device results, IN_DEVICE|OUT_DEVICE all on the device."""
        check(lib().pgb_orb_extract(self._h, gray_ptr, int(where), n, w, h, pitch, frame_stride, kps_ptr, desc_ptr,
                                    counts_ptr, cap))

    def check(self):
        check(lib().pgb_orb_check(self._h))

    def run_stage(self, which: int):
        check(lib().pgb_orb_run_stage(self._h, which))

    # ---- mvImagePyramid and stage products of the last call
    def _get_img(self, fn, frame, level):
        w = C.c_int(); h = C.c_int()
        check(fn(self._h, frame, level, None, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        check(fn(self._h, frame, level, np_ptr(out), C.byref(w), C.byref(h)))
        return out

    def image_pyramid(self, level: int, frame: int = 0) -> np.ndarray:
        return self._get_img(lib().pgb_orb_get_level, frame, level)

    def score_map(self, level: int, frame: int = 0) -> np.ndarray:
        return self._get_img(lib().pgb_orb_get_score_map, frame, level)

    def blurred_level(self, level: int, frame: int = 0) -> np.ndarray:
        return self._get_img(lib().pgb_orb_get_blurred_level, frame, level)

    def candidates(self, level: int, frame: int = 0) -> np.ndarray:
        n = C.c_int32()
        check(lib().pgb_orb_get_candidates(self._h, frame, level, None, 0, C.byref(n)))
        out = np.empty((max(n.value, 1), 3), np.int32)
        check(lib().pgb_orb_get_candidates(self._h, frame, level, np_ptr(out), n.value, C.byref(n)))
        return out[:n.value]

    def level_size(self, w: int, h: int, level: int):
        lw = C.c_int(); lh = C.c_int()
        check(lib().pgb_orb_level_size(self._h, w, h, level, C.byref(lw), C.byref(lh)))
        return lw.value, lh.value


def frames_to_gray_rotated(frames: np.ndarray, rotate_degrees: int, rgb_order: bool = True, vertical_flip: bool = False,
                           horizontal_flip: bool = False, formula: int = 0, device: int = 0) -> np.ndarray:
    """The frame feed after the decoder: the container's `rotate` metadata (image_sequence_reader.cc:113-118, 186-207), then
    cv::flip, then cvtColor to gray, of (n, h, w[, c]) uint8 frames on the device.  Returns (n, w, h) for 90 / 270."""
    a = np.ascontiguousarray(frames, np.uint8)
    if a.ndim == 3:
        a = a[..., None]
    n, h, w, c = a.shape
    oh, ow = (w, h) if rotate_degrees % 360 in (90, 270) else (h, w)
    out = np.empty((n, oh, ow), np.uint8)
    check(lib().pgb_frames_to_gray_rotated(device, np_ptr(a), 0, n, w, h, c, int(rgb_order), w * c, w * c * h, int(rotate_degrees),
                                           int(vertical_flip), int(horizontal_flip), formula, np_ptr(out), 0, ow, ow * oh, None))
    return out


def frames_to_gray(frames: np.ndarray, rgb_order: bool = True, vertical_flip: bool = False, horizontal_flip: bool = False,
                   formula: int = 0, device: int = 0) -> np.ndarray:
    """cv::flip + cvtColor to gray of (n, h, w[, c]) uint8 frames on the device (image_sequence_reader.cc:163-175,
    Tracking.cc:243-258).  formula 0 = OpenCV 2.4 fixed point (the reference's pinned version), 1 = OpenCV >= 3."""
    a = np.ascontiguousarray(frames, np.uint8)
    if a.ndim == 3:
        a = a[..., None]
    n, h, w, c = a.shape
    out = np.empty((n, h, w), np.uint8)
    check(lib().pgb_frames_to_gray(device, np_ptr(a), 0, n, w, h, c, int(rgb_order), w * c, w * c * h, int(vertical_flip),
                                   int(horizontal_flip), formula, np_ptr(out), 0, w, w * h, None))
    return out
