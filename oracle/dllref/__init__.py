"""ctypes binding of oracle/_ref/libgms_dll_host.so — the reference's own matchGMS machine code, hosted on Linux.

TEST INFRASTRUCTURE ONLY (see gms_dll_host.c).  Usable only where /root/reference exists (the build container):
tests that need it skip otherwise and rely on the committed outputs tests/golden/gms_dll_*.npz
(generator: tests/golden/make_gms_dll_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE = os.path.dirname(_HERE)
DLL_PATH = "/root/reference/SfM-GMS/bin/opencv_xfeatures2d452.dll"
_LIB = None

KEYPOINT_DT = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                        ("octave", "<i4"), ("class_id", "<i4")])          # cv::KeyPoint, 28 bytes
DMATCH_DT = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])  # cv::DMatch


def available():
    return os.path.exists(DLL_PATH)


def build(force=False):
    so = os.path.join(_ORACLE, "_ref", "libgms_dll_host.so")
    src = os.path.join(_HERE, "gms_dll_host.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _ORACLE, "ref"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        if not available():
            raise RuntimeError("the reference DLL is not present on this machine: " + DLL_PATH)
        L = ctypes.CDLL(build())
        L.gmsdll_error.restype = ctypes.c_char_p
        L.gmsdll_load.argtypes = [ctypes.c_char_p]
        vp, i64, ci, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
        L.gmsdll_tables.argtypes = [vp, vp]
        L.gmsdll_grid_left.argtypes = [vp, i64, ci, vp]
        L.gmsdll_grid_right.argtypes = [vp, i64, ci, ci, vp]
        L.gmsdll_match_gms.argtypes = [ci, ci, ci, ci, vp, i64, vp, i64, vp, i64, ci, ci, dbl, vp]
        L.gmsdll_match_gms.restype = i64
        L.gmsdll_hypotheses.argtypes = [ci, ci, ci, ci, vp, i64, vp, i64, vp, i64, dbl, vp, vp]
        if L.gmsdll_load(DLL_PATH.encode()) != 0:
            raise RuntimeError("gmsdll_load: " + L.gmsdll_error().decode())
        _LIB = L
    return _LIB


def tables():
    """-> (ROT int32[8][9] 1-based, SCALE float64[5]) as they sit in the mapped image after its static initialiser."""
    rot = np.zeros(72, np.int32)
    sc = np.zeros(5, np.float64)
    lib().gmsdll_tables(rot.ctypes.data, sc.ctypes.data)
    return rot.reshape(8, 9), sc


def grid_left(norm_pts, gtype):
    """GMSMatcher::getGridIndexLeft (@VA 0x180047bc0) on normalised points [n,2] f32, type 1..4."""
    p = np.ascontiguousarray(norm_pts, np.float32).reshape(-1, 2)
    out = np.empty(len(p), np.int32)
    lib().gmsdll_grid_left(p.ctypes.data, len(p), int(gtype), out.ctypes.data)
    return out


def grid_right(norm_pts, wr, hr):
    """GMSMatcher::getGridIndexRight (@VA 0x180047d60) with mGridSizeRight = (wr, hr)."""
    p = np.ascontiguousarray(norm_pts, np.float32).reshape(-1, 2)
    out = np.empty(len(p), np.int32)
    lib().gmsdll_grid_right(p.ctypes.data, len(p), int(wr), int(hr), out.ctypes.data)
    return out


def _kp(xy):
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    k = np.zeros(len(xy), KEYPOINT_DT)
    k["x"], k["y"] = xy[:, 0], xy[:, 1]
    k["size"], k["angle"], k["class_id"] = 31.0, -1.0, -1
    return k


def _dm(query_idx, train_idx, dist=None, tag_positions=False):
    m = np.zeros(len(query_idx), DMATCH_DT)
    m["queryIdx"], m["trainIdx"] = query_idx, train_idx
    m["distance"] = 0 if dist is None else dist
    if tag_positions:            # GMS never reads imgIdx; the DLL copies whole records, so the tag survives
        m["imgIdx"] = np.arange(len(m))
    return m


def match_gms(size1, size2, kp1_xy, kp2_xy, query_idx, train_idx, with_rotation=False, with_scale=False,
              threshold_factor=6.0, dist=None, tag_positions=False):
    """The DLL's exported cv::xfeatures2d::matchGMS (@VA 0x180048280).  size = (width, height).
    -> matchesGMS as a DMATCH_DT array (in input order, as the reference returns it)."""
    k1, k2, m = _kp(kp1_xy), _kp(kp2_xy), _dm(query_idx, train_idx, dist, tag_positions)
    out = np.zeros(max(len(m), 1), DMATCH_DT)
    n = lib().gmsdll_match_gms(int(size1[0]), int(size1[1]), int(size2[0]), int(size2[1]), k1.ctypes.data, len(k1),
                               k2.ctypes.data, len(k2), m.ctypes.data, len(m), int(bool(with_rotation)),
                               int(bool(with_scale)), float(threshold_factor), out.ctypes.data)
    return out[:n]


def hypotheses(size1, size2, kp1_xy, kp2_xy, query_idx, train_idx, threshold_factor=6.0, want_masks=False):
    """GMSMatcher ctor + setScale(s) + run(rot) for all 5 x 8 hypotheses -> counts int32[40] (scale-major)
    and, optionally, the 40 inlier masks [40, n] bool."""
    k1, k2, m = _kp(kp1_xy), _kp(kp2_xy), _dm(query_idx, train_idx)
    counts = np.zeros(40, np.int32)
    masks = np.zeros((40, len(m)), np.uint8) if want_masks else None
    rc = lib().gmsdll_hypotheses(int(size1[0]), int(size1[1]), int(size2[0]), int(size2[1]), k1.ctypes.data, len(k1),
                                 k2.ctypes.data, len(k2), m.ctypes.data, len(m), float(threshold_factor),
                                 counts.ctypes.data, masks.ctypes.data if want_masks else None)
    if rc:
        raise RuntimeError("gmsdll_hypotheses: " + lib().gmsdll_error().decode())
    return (counts, masks.astype(bool)) if want_masks else counts
