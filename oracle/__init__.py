"""ctypes loader for the CPU oracle (oracle/sfmgms_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from sfm_gms_b200/ (the product path).
Reference behaviour restated: FeatureMatchUtil.cpp:66-69 (BFMatcher::match + matchGMS); see the
C file's header for the pinning status.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libsfmgms_oracle.so")
    src = os.path.join(_HERE, "sfmgms_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libsfmgms_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        L.oracle_bf_hamming.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, i32p, i32p, ip]
        L.oracle_bf_hamming.restype = ctypes.c_int
        L.oracle_bf_hamming_crosscheck.argtypes = [u8p, ctypes.c_int, u8p, ctypes.c_int, ctypes.c_int, i32p, i32p, u8p]
        L.oracle_bf_hamming_crosscheck.restype = ctypes.c_int
        L.oracle_gms.argtypes = [ctypes.c_int] * 4 + [f32p, ctypes.c_int, ctypes.c_int, f32p, ctypes.c_int,
                                                      ctypes.c_int, i32p, i32p, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_int, ctypes.c_int, ctypes.c_double, u8p, ip, ip,
                                                      ip, ip]
        L.oracle_gms.restype = ctypes.c_int
        L.oracle_gms_ex.argtypes = L.oracle_gms.argtypes + [u8p]
        L.oracle_gms_ex.restype = ctypes.c_int
        L.oracle_gms_tables.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_gms_grid_left.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]
        L.oracle_gms_grid_right.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.oracle_gms_right_grid.argtypes = [ctypes.c_int]
        L.oracle_bf_l2.argtypes = [f32p, ctypes.c_int, f32p, ctypes.c_int, ctypes.c_int, i32p, f32p, ip]
        L.oracle_bf_l2.restype = ctypes.c_int
        L.oracle_orb_blur7.argtypes = [u8p, ctypes.c_int, ctypes.c_int, u8p]
        L.oracle_orb_blur7.restype = ctypes.c_int
        L.oracle_set_num_threads.argtypes = [ctypes.c_int]
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def bf_hamming(q, t):
    """-> (train_idx int32[nq], dist int32[nq]); both empty if t is empty (cv2 returns no matches)."""
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    nq, nt = q.shape[0], t.shape[0]
    db = q.shape[1] if q.ndim == 2 else t.shape[1]
    idx = np.empty(nq, np.int32)
    dist = np.empty(nq, np.int32)
    n = ctypes.c_int(0)
    rc = lib().oracle_bf_hamming(_p(q, ctypes.c_uint8), nq, _p(t, ctypes.c_uint8), nt, db,
                                 _p(idx, ctypes.c_int32), _p(dist, ctypes.c_int32), ctypes.byref(n))
    if rc:
        raise ValueError("oracle_bf_hamming rc=%d" % rc)
    return idx[: n.value], dist[: n.value]


def bf_l2(q, t):
    """cv2.BFMatcher(cv2.NORM_L2).match semantics -> (train_idx int32[n], dist float32[n])."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.empty(nq, np.int32)
    dist = np.empty(nq, np.float32)
    n = ctypes.c_int(0)
    rc = lib().oracle_bf_l2(_p(q, ctypes.c_float), nq, _p(t, ctypes.c_float), nt, q.shape[1], _p(idx, ctypes.c_int32),
                            _p(dist, ctypes.c_float), ctypes.byref(n))
    if rc:
        raise ValueError("oracle_bf_l2 rc=%d" % rc)
    return idx[: n.value], dist[: n.value]


def bf_hamming_crosscheck(q, t):
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    nq, nt = q.shape[0], t.shape[0]
    idx = np.empty(nq, np.int32)
    dist = np.empty(nq, np.int32)
    keep = np.zeros(nq, np.uint8)
    rc = lib().oracle_bf_hamming_crosscheck(_p(q, ctypes.c_uint8), nq, _p(t, ctypes.c_uint8), nt, q.shape[1],
                                            _p(idx, ctypes.c_int32), _p(dist, ctypes.c_int32),
                                            _p(keep, ctypes.c_uint8))
    if rc:
        raise ValueError("oracle_bf_hamming_crosscheck rc=%d" % rc)
    return idx, dist, keep.astype(bool)


def orb_blur7(gray):
    """ORB's 7x7 sigma-2 Gaussian of one pyramid level (float32 separable filter with OpenCV's FMA order; C file)."""
    g = np.ascontiguousarray(gray, dtype=np.uint8)
    out = np.empty_like(g)
    rc = lib().oracle_orb_blur7(_p(g, ctypes.c_uint8), g.shape[1], g.shape[0], _p(out, ctypes.c_uint8))
    if rc:
        raise ValueError("oracle_orb_blur7 rc=%d" % rc)
    return out


def bf_l2_crosscheck(q, t):
    """cv2.BFMatcher(cv2.NORM_L2, crossCheck=True).match (FeatureMatchUtil.cpp:22-23): the forward nearest
    neighbour of every query is kept iff the nearest QUERY of that train row (lowest index on ties) is the
    query itself.  -> (train_idx, dist float32, keep bool), all of length nq."""
    idx, dist = bf_l2(q, t)
    if len(idx) == 0:
        return idx, dist, np.zeros(0, bool)
    ridx, _ = bf_l2(t, q)
    return idx, dist, ridx[idx] == np.arange(len(idx))


def brute_force_match(q, t, norm="l2", cross_check=True, distance_coef=4.0, max_matching_size=500):
    """bruteForceMatch (FeatureMatchUtil.cpp:20-31) / match (:38-50): (cross-checked) matches, sorted by
    distance (stable: equal distances stay in queryIdx order -- std::sort leaves that order unspecified),
    pruned while front*coef < back (double arithmetic on float distances), capped.
    -> (query_idx int32, train_idx int32, dist float32)."""
    if norm == "l2":
        idx, dist, keep = bf_l2_crosscheck(q, t) if cross_check else bf_l2(q, t) + (None,)
    else:
        idx, dist, keep = bf_hamming_crosscheck(q, t) if cross_check else bf_hamming(q, t) + (None,)
    qi = np.arange(len(idx), dtype=np.int32)
    if keep is not None:
        qi, idx, dist = qi[keep], idx[keep], dist[keep]
    dist = dist.astype(np.float32)
    order = np.argsort(dist, kind="stable")
    qi, idx, dist = qi[order], idx[order], dist[order]
    n = len(dist)
    while n > 0 and np.float64(dist[0]) * distance_coef < np.float64(dist[n - 1]):
        n -= 1
    n = min(n, max_matching_size)
    return qi[:n], idx[:n].astype(np.int32), dist[:n]


def gms_tables():
    """-> (ROT int32[8][9], SCALE float64[5]) of the restatement (compared with the DLL's in tests/test_gms_dll.py)."""
    rot = np.zeros(72, np.int32)
    sc = np.zeros(5, np.float64)
    lib().oracle_gms_tables(rot.ctypes.data, sc.ctypes.data)
    return rot.reshape(8, 9), sc


def gms_right_grid(s):
    return lib().oracle_gms_right_grid(int(s))


def gms_grid_left(norm_pts, gtype):
    p = np.ascontiguousarray(norm_pts, np.float32).reshape(-1, 2)
    out = np.empty(len(p), np.int32)
    lib().oracle_gms_grid_left(p.ctypes.data, len(p), int(gtype), out.ctypes.data)
    return out


def gms_grid_right(norm_pts, wr, hr):
    p = np.ascontiguousarray(norm_pts, np.float32).reshape(-1, 2)
    out = np.empty(len(p), np.int32)
    lib().oracle_gms_grid_right(p.ctypes.data, len(p), int(wr), int(hr), out.ctypes.data)
    return out


def gms(size1, size2, kp1_xy, kp2_xy, query_idx, train_idx, with_rotation=False, with_scale=False,
        threshold_factor=6.0, want_all_masks=False):
    """matchGMS semantics.  size = (width, height).  Returns dict(mask, n_inliers, hyp_counts, best_hyp).

    mask has length n_matches, or 0 when rotation/scale search found nothing (reference quirk).
    want_all_masks: also 'all_masks' bool[40, n] — the mask of every hypothesis that was run.
    """
    kp1 = np.ascontiguousarray(kp1_xy, dtype=np.float32).reshape(-1, 2)
    kp2 = np.ascontiguousarray(kp2_xy, dtype=np.float32).reshape(-1, 2)
    qi = np.ascontiguousarray(query_idx, dtype=np.int32)
    ti = np.ascontiguousarray(train_idx, dtype=np.int32)
    n = qi.shape[0]
    mask = np.zeros(max(n, 1), np.uint8)
    mlen, ninl, bh = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(-1)
    hyp = np.full(40, -1, np.int32)
    allm = np.zeros((40, max(n, 1)), np.uint8) if want_all_masks else None
    rc = lib().oracle_gms_ex(int(size1[0]), int(size1[1]), int(size2[0]), int(size2[1]),
                          _p(kp1, ctypes.c_float), kp1.shape[0], 2, _p(kp2, ctypes.c_float), kp2.shape[0], 2,
                          _p(qi, ctypes.c_int32), _p(ti, ctypes.c_int32), 1, n, int(bool(with_rotation)),
                          int(bool(with_scale)), float(threshold_factor), _p(mask, ctypes.c_uint8),
                          ctypes.byref(mlen), ctypes.byref(ninl), _p(hyp, ctypes.c_int), ctypes.byref(bh),
                          _p(allm, ctypes.c_uint8) if want_all_masks else None)
    if rc:
        raise ValueError("oracle_gms rc=%d" % rc)
    out = dict(mask=mask[: mlen.value].astype(bool), n_inliers=ninl.value, hyp_counts=hyp, best_hyp=bh.value)
    if want_all_masks:
        out["all_masks"] = allm[:, :n].astype(bool)
    return out
