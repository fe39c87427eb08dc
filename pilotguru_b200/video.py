"""Host-side mirror of the reference's video frame source for the files this build decodes.

`VideoImageSequenceSource` (src/io/image_sequence_reader.cc:74-208, interface include/io/image_sequence_reader.hpp)
hands the tracker `TimestampedImage{image (RGB24, rotated), timestamp, frame_id}` one frame at a time through
`hasNext()` / `next()`.  Here the container is demuxed by `pgb_video_open` and the frames are decoded by nvJPEG into
device memory (`pgb_video_read_rgb`); `next()` returns host copies for parity tests, `read_gray_device()` is the path the
extractor uses (decode -> rotate / flip -> gray, all on the device).  Motion-JPEG AVI only: any other codec raises with
its FOURCC in the message.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import PgbError, check, last_error, lib, np_ptr


class VideoImageSequenceSource:
    def __init__(self, filename: str, device: int = 0):
        self._h = lib().pgb_video_open(device, filename.encode())
        if not self._h:
            raise PgbError(-1, last_error())
        w, h, rot = C.c_int(), C.c_int(), C.c_int()
        n, fps = C.c_int64(), C.c_double()
        check(lib().pgb_video_info(self._h, C.byref(w), C.byref(h), C.byref(n), C.byref(fps), C.byref(rot)))
        self.width, self.height, self.n_frames, self.fps, self.rotate_degrees = w.value, h.value, n.value, fps.value, rot.value
        self.device = device
        self._next = 0

    def close(self):
        if self._h:
            lib().pgb_video_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def frame_span(self, i: int):
        off, size = C.c_uint64(), C.c_uint32()
        check(lib().pgb_video_frame_span(self._h, i, C.byref(off), C.byref(size)))
        return off.value, size.value

    def hasNext(self) -> bool:
        return self._next < self.n_frames

    def read_rgb(self, first: int, n: int):
        """Frames [first, first + n) as a host (n, h, w, 3) RGB array + their timestamps in seconds (decoded on the device)."""
        import torch
        dev = torch.empty((n, self.height, self.width, 3), dtype=torch.uint8, device=f"cuda:{self.device}")
        ts = np.zeros(n, np.float64)
        st = torch.cuda.current_stream(self.device)
        check(lib().pgb_video_read_rgb(self._h, first, n, dev.data_ptr(), self.width * 3, self.width * self.height * 3, np_ptr(ts),
                                       st.cuda_stream))
        st.synchronize()
        return dev.cpu().numpy(), ts

    def read_gray_device(self, first: int, n: int, rgb_order: bool = True, vertical_flip: bool = False, horizontal_flip: bool = False,
                         formula: int = 0):
        """Decode -> rotate -> flip -> gray on the device: a torch uint8 (n, H, W) tensor ready for extract_ptr, + timestamps."""
        import torch
        dev = torch.empty((n, self.height, self.width, 3), dtype=torch.uint8, device=f"cuda:{self.device}")
        ts = np.zeros(n, np.float64)
        st = torch.cuda.current_stream(self.device)
        check(lib().pgb_video_read_rgb(self._h, first, n, dev.data_ptr(), self.width * 3, self.width * self.height * 3, np_ptr(ts),
                                       st.cuda_stream))
        oh, ow = (self.width, self.height) if self.rotate_degrees in (90, 270) else (self.height, self.width)
        gray = torch.empty((n, oh, ow), dtype=torch.uint8, device=dev.device)
        check(lib().pgb_frames_to_gray_rotated(self.device, dev.data_ptr(), 1, n, self.width, self.height, 3, int(rgb_order), self.width * 3,
                                               self.width * self.height * 3, self.rotate_degrees, int(vertical_flip), int(horizontal_flip),
                                               formula, gray.data_ptr(), 1, ow, ow * oh, st.cuda_stream))
        st.synchronize()
        return gray, ts

    def next(self):
        """(image RGB24 (h, w, 3) uint8, timestamp in seconds, frame_id) -- frame_id counts from 1 like the reference's
        pre-incremented next_frame_.frame_id (:172)."""
        if not self.hasNext():
            raise PgbError(-1, "CHECK failed: has_next_")   # CHECK(has_next_) (:134)
        img, ts = self.read_rgb(self._next, 1)
        self._next += 1
        return img[0], float(ts[0]), self._next
