"""GPU tests of the reference-side bindings (integration/*.cc, INTEGRATION.md sections 1-2) as compiled code: the
reference's OWN class ORB_SLAM2::ORBextractor (thirdparty/orb-slam2/include/ORBextractor.h, unmodified) with the adapter's
replacement bodies, and ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) / DescriptorDistance behind the
reference's member names, both calling libpgb200.so -- results identical to the oracle's (which is pinned to the
reference's own ORBextractor.cc / ORBmatcher.cc bodies in tests/test_oracle_reference_pin.py).
oracle/_ref/libpgb_adapters.so is built by `make -C oracle _ref` where /root/reference exists and travels to the GPU box."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from pilotguru_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libpgb_adapters.so")


@pytest.fixture(scope="module")
def ada():
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "adapters"], check=True, capture_output=True)
    if not os.path.exists(SO):
        pytest.skip("oracle/_ref/libpgb_adapters.so is not built and /root/reference is absent")
    l = C.CDLL(SO)
    l.pga_orb_create.restype = C.c_void_p
    return l


def _extract(l, h, img, cap=4096):
    kps = np.zeros((cap, 7), np.float32); desc = np.zeros((cap, 32), np.uint8)
    n = l.pga_orb_extract(C.c_void_p(h), img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], kps.ctypes.data_as(C.c_void_p),
                          desc.ctypes.data_as(C.c_void_p), cap)
    return kps[:n], desc[:n]


def test_orbextractor_class_with_adapter_bodies_equals_the_oracle(ada):
    nf = 1000
    h = ada.pga_orb_create(nf, C.c_float(1.2), 8, 20, 7)
    orc = O.OrbOracle(nf, 1.2, 8, 20, 7)
    t = orc.tables()
    sc, inv, s2, is2 = (np.zeros(8, np.float32) for _ in range(4))
    ada.pga_orb_tables(C.c_void_p(h), *(a.ctypes.data_as(C.c_void_p) for a in (sc, inv, s2, is2)))
    assert np.array_equal(sc, t["scale"]) and np.array_equal(inv, t["inv_scale"]) and np.array_equal(s2, t["sigma2"]) and np.array_equal(is2, t["inv_sigma2"])
    for img in (synth.frame(5), synth.frame(2, w=640, h=480), synth.frame(1, w=701, h=403)):
        img = np.ascontiguousarray(img)
        k, d = _extract(ada, h, img)
        ok, od = orc.extract(img)
        assert len(k) == len(ok) > 300
        for i, f in enumerate(["x", "y", "size", "angle", "response"]):
            assert np.array_equal(k[:, i].view(np.uint32), ok[f].view(np.uint32)), f
        assert np.array_equal(k[:, 5].astype(np.int32), ok["octave"]) and (k[:, 6] == -1).all()
        assert np.array_equal(d, od)
        w, hh = C.c_int(), C.c_int()
        ada.pga_orb_level(C.c_void_p(h), 3, None, C.byref(w), C.byref(hh))
        lvl = np.zeros((hh.value, w.value), np.uint8)
        ada.pga_orb_level(C.c_void_p(h), 3, lvl.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(hh))
        assert np.array_equal(lvl, orc.level(3))                              # the public mvImagePyramid member
    k, d = _extract(ada, h, np.full((480, 640), 77, np.uint8))               # flat image: no keypoints, descriptors released
    assert len(k) == 0
    ada.pga_orb_destroy(C.c_void_p(h))


@pytest.mark.parametrize("check_ori", [1, 0])
def test_orbmatcher_search_by_projection_adapter_equals_the_oracle(ada, check_ori):
    orc = O.OrbOracle(1000, 1.2, 8, 20, 7)
    sf = orc.tables()["scale"]
    (k0, d0), (k1, d1) = orc.extract(synth.frame(3)), orc.extract(synth.frame(4))
    fl = synth.flow(4)
    uv = np.stack([k0["x"] + np.float32(fl[0]), k0["y"] + np.float32(fl[1])], axis=1).astype(np.float32)
    valid = np.ones(len(k0), np.uint8); valid[::5] = 0
    for th in (15.0, 30.0):
        on, om, _ = O.search_by_projection(k1, d1, uv, k0["octave"], k0["angle"], d0, valid, (0, 1920, 0, 1080), th, sf, bool(check_ori))
        m = np.full(len(k1), -9, np.int32)
        c = lambda a, t: np.ascontiguousarray(a, t)
        n = ada.pga_search_by_projection(c(k1, O.KP_DTYPE).ctypes.data_as(C.c_void_p), c(d1, np.uint8).ctypes.data_as(C.c_void_p), len(k1),
                                         uv.ctypes.data_as(C.c_void_p), c(k0["octave"], np.int32).ctypes.data_as(C.c_void_p),
                                         c(k0["angle"], np.float32).ctypes.data_as(C.c_void_p), c(d0, np.uint8).ctypes.data_as(C.c_void_p),
                                         valid.ctypes.data_as(C.c_void_p), len(k0), C.c_float(0), C.c_float(1920), C.c_float(0), C.c_float(1080),
                                         C.c_float(th), c(sf, np.float32).ctypes.data_as(C.c_void_p), len(sf), check_ori, m.ctypes.data_as(C.c_void_p))
        assert n == on > 300 and np.array_equal(m, om)
    a, b = d0[0].copy(), d1[0].copy()
    assert ada.pga_descriptor_distance(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)) == int(np.unpackbits(a ^ b).sum())
