"""Host-side mirror of the reference's matching interface, over the C ABI (include/sfmgms.h).

Mirrors, with the same names / argument meaning / error behaviour:
  * ``cv::BFMatcher(NORM_HAMMING[, crossCheck])::match(query, train)``   FeatureMatchUtil.cpp:22-23, 66-68
  * ``cv::xfeatures2d::matchGMS(size1, size2, kp1, kp2, matches1to2, withRotation, withScale,
    thresholdFactor)``                                                    FeatureMatchUtil.cpp:69
  * upstream ``gms_matcher(vkp1, size1, vkp2, size2, vDMatches).GetInlierMask(vbInliers, WithScale,
    WithRotation)`` (scale BEFORE rotation — SURVEY fact 4).
Everything here is plumbing (ctypes, numpy views); all arithmetic runs in libsfmgms.so on the GPU.
"""
import ctypes
import os
import threading
from collections import namedtuple

import numpy as np

NORM_HAMMING = 6  # cv::NORM_HAMMING
NORM_L2 = 4       # cv::NORM_L2

SFMGMS_HOST, SFMGMS_DEVICE = 0, 1
OPT_HAMMING_KERNEL, OPT_GMS_CHUNK_BYTES, OPT_TIMING, OPT_TC_OPERAND_CACHE, OPT_L2_KERNEL = 1, 2, 3, 4, 5
OPT_CHUNK_ROWS, OPT_GMS_DENSE, OPT_OVERLAP, OPT_COMPACT_RECORD = 6, 7, 8, 9
HAMMING_AUTO, HAMMING_POPC, HAMMING_TC, HAMMING_FP4 = 0, 1, 2, 3

_ERR_NAMES = {1: "ERR_ARG", 2: "ERR_TRAIN_ROWS", 3: "ERR_DOMAIN", 4: "ERR_INDEX", 5: "ERR_CUDA", 6: "ERR_STATE",
              7: "ERR_CAPACITY", 8: "ERR_NCCL"}

# cv::DMatch, 16 bytes (the record sfmgms_match_pairs_compact writes)
DMATCH_DT = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])

DMatch = namedtuple("DMatch", ["queryIdx", "trainIdx", "imgIdx", "distance"])


class SfmGmsError(RuntimeError):
    """Raised where OpenCV would throw cv::Exception (and for CUDA failures)."""

    def __init__(self, code, msg):
        super().__init__("sfmgms %s (%d): %s" % (_ERR_NAMES.get(code, "ERR"), code, msg))
        self.code = code


_LIB = None
# SFMGMS_LIB: developer override to load an experimental build of the same library (kernel tuning only)
_LIB_PATH = os.environ.get("SFMGMS_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsfmgms.so")


def load_library():
    """Load libsfmgms.so.  Fails loudly: there is no fallback implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m sfm_gms_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback." % _LIB_PATH)
    L = ctypes.CDLL(_LIB_PATH)
    c_int, c_void_p, c_double, c_i64 = ctypes.c_int, ctypes.c_void_p, ctypes.c_double, ctypes.c_int64
    P = ctypes.POINTER
    L.sfmgms_create.argtypes = [P(c_void_p), c_int]
    L.sfmgms_destroy.argtypes = [c_void_p]
    L.sfmgms_destroy.restype = None
    L.sfmgms_last_error.argtypes = [c_void_p]
    L.sfmgms_last_error.restype = ctypes.c_char_p
    L.sfmgms_version.restype = c_int
    L.sfmgms_set_option.argtypes = [c_void_p, c_int, c_i64]
    L.sfmgms_kernel_launches.argtypes = [c_void_p]
    L.sfmgms_kernel_launches.restype = c_i64
    L.sfmgms_last_timing.argtypes = [c_void_p, c_void_p]
    L.sfmgms_stream.argtypes = [c_void_p]
    L.sfmgms_stream.restype = c_void_p
    L.sfmgms_bf_hamming.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, P(c_int)]
    L.sfmgms_bf_l2.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, P(c_int)]
    L.sfmgms_bf_hamming_crosscheck.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                               c_void_p, c_void_p]
    L.sfmgms_bf_l2_crosscheck.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p]
    L.sfmgms_brute_force_match.argtypes = [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_double,
                                           c_int, c_void_p, c_void_p, c_void_p, c_int, P(c_int)]
    L.sfmgms_orb_compute.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_void_p, c_void_p, P(c_int)]
    L.sfmgms_orb_detect_and_compute.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                                c_void_p, c_int, P(c_int)]
    L.sfmgms_orb_detect_and_compute_ex.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                                   c_void_p, c_int, P(c_int)]
    L.sfmgms_set_images_from_pixels.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p]
    L.sfmgms_get_image_keypoints.argtypes = [c_void_p, c_int, c_void_p, c_int, P(c_int)]
    L.sfmgms_gms.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                             c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_void_p, P(c_int),
                             P(c_int), P(c_int)]
    L.sfmgms_match_pair.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_void_p, c_void_p,
                                    c_void_p, P(c_int), P(c_int), P(c_int)]
    L.sfmgms_set_images.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
    L.sfmgms_match_offsets.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
    L.sfmgms_match_pairs.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]
    L.sfmgms_match_image_set.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                         c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.sfmgms_inlier_points.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, P(c_int)]
    L.sfmgms_match_pairs_compact.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_i64, P(c_i64)]
    L.sfmgms_match_pairs_async.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p]
    L.sfmgms_match_pairs_compact_async.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_i64]
    L.sfmgms_wait.argtypes = [c_void_p, P(c_i64)]
    L.sfmgms_gms_hypotheses.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                                        c_void_p, c_void_p, c_int, c_int, c_double, c_void_p]
    L.sfmgms_multi_create.argtypes = [P(c_void_p), c_void_p, c_int]
    L.sfmgms_multi_destroy.argtypes = [c_void_p]
    L.sfmgms_multi_destroy.restype = None
    L.sfmgms_multi_last_error.argtypes = [c_void_p]
    L.sfmgms_multi_last_error.restype = ctypes.c_char_p
    L.sfmgms_multi_device_count.argtypes = [c_void_p]
    L.sfmgms_multi_context.argtypes = [c_void_p, c_int]
    L.sfmgms_multi_context.restype = c_void_p
    L.sfmgms_multi_last_broadcast_ms.argtypes = [c_void_p]
    L.sfmgms_multi_last_broadcast_ms.restype = c_double
    L.sfmgms_multi_set_images.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    L.sfmgms_multi_match_pairs.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p]
    L.sfmgms_multi_match_pairs_compact.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_i64, P(c_i64)]
    L.sfmgms_kernel_times.argtypes = [c_void_p, ctypes.c_char_p, c_int]
    L.sfmgms_device_bytes.argtypes = [c_void_p]
    L.sfmgms_device_bytes.restype = c_i64
    L.sfmgms_host_alloc.argtypes = [ctypes.c_size_t, P(c_void_p)]
    L.sfmgms_host_free.argtypes = [c_void_p]
    _LIB = L
    return L


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


class _HostBlock:
    """One sfmgms_host_alloc allocation; freed when the last array viewing it goes away."""

    def __init__(self, nbytes):
        self._lib = load_library()
        p = ctypes.c_void_p()
        rc = self._lib.sfmgms_host_alloc(nbytes, ctypes.byref(p))
        if rc != 0:
            raise SfmGmsError(rc, "sfmgms_host_alloc(%d bytes) failed" % nbytes)
        self.ptr = p.value or 0
        self.buf = (ctypes.c_uint8 * nbytes).from_address(self.ptr) if nbytes else (ctypes.c_uint8 * 0)()
        self.buf._owner = self     # the ctypes view keeps the block alive for as long as numpy holds it

    def __del__(self):
        if getattr(self, "ptr", 0):
            self._lib.sfmgms_host_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


def host_empty(shape, dtype):
    """numpy array in page-locked host memory from sfmgms_host_alloc (contents undefined)."""
    dt = np.dtype(dtype)
    shape = (shape,) if np.isscalar(shape) else tuple(shape)
    n = int(np.prod(shape, dtype=np.int64)) * dt.itemsize
    blk = _HostBlock(n)
    return np.frombuffer(blk.buf, dtype=dt, count=n // dt.itemsize).reshape(shape)


def host_zeros(shape, dtype):
    a = host_empty(shape, dtype)
    a[...] = 0
    return a


def host_array(a):
    """Copy of `a` (C order) in page-locked host memory."""
    a = np.asarray(a)
    out = host_empty(a.shape, a.dtype)
    out[...] = a
    return out


def _kp_xy(kps):
    """cv2.KeyPoint sequence or (N,2) array -> C-contiguous float32 (N,2)."""
    if isinstance(kps, np.ndarray):
        a = np.ascontiguousarray(kps, dtype=np.float32)
        return a.reshape(-1, 2)
    if len(kps) == 0:
        return np.zeros((0, 2), np.float32)
    return np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2)


def _match_idx(matches):
    """Sequence of DMatch-like objects, or (queryIdx, trainIdx) arrays -> two int32 arrays."""
    if isinstance(matches, tuple) and len(matches) == 2 and isinstance(matches[0], np.ndarray):
        return (np.ascontiguousarray(matches[0], dtype=np.int32), np.ascontiguousarray(matches[1], dtype=np.int32))
    n = len(matches)
    q = np.fromiter((m.queryIdx for m in matches), dtype=np.int32, count=n)
    t = np.fromiter((m.trainIdx for m in matches), dtype=np.int32, count=n)
    return q, t


def _size(sz):
    """cv::Size convention: (width, height)."""
    return int(sz[0]), int(sz[1])


class Context:
    """One per host thread / per GPU (SURVEY §8b threading)."""

    def __init__(self, device=0):
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.sfmgms_create(ctypes.byref(h), int(device))
        if rc:
            raise SfmGmsError(rc, self._lib.sfmgms_last_error(None).decode())
        self._h = h
        self.device = int(device)
        self._keep = None  # keeps adopted device tensors alive

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sfmgms_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SfmGmsError(rc, self._lib.sfmgms_last_error(self._h).decode())

    # -- options / introspection -------------------------------------------------------------------
    def set_option(self, key, value):
        self._check(self._lib.sfmgms_set_option(self._h, int(key), int(value)))

    @property
    def kernel_launches(self):
        return int(self._lib.sfmgms_kernel_launches(self._h))

    def last_timing(self):
        """(hamming_ms, gms_ms, hamming_launches) of the last batch; needs set_option(OPT_TIMING, 1)."""
        out = np.zeros(3, np.float64)
        self._check(self._lib.sfmgms_last_timing(self._h, _ptr(out)))
        return float(out[0]), float(out[1]), int(out[2])

    @property
    def stream(self):
        return int(self._lib.sfmgms_stream(self._h) or 0)

    # -- stage 1 ------------------------------------------------------------------------------------
    def bf_hamming(self, query, train):
        """-> (train_idx int32[n], dist int32[n]); n = len(query), or 0 if train is empty."""
        q, t = self._desc(query), self._desc(train)
        idx = np.empty(q.shape[0], np.int32)
        dist = np.empty(q.shape[0], np.int32)
        n = ctypes.c_int(0)
        self._check(self._lib.sfmgms_bf_hamming(self._h, _ptr(q), q.shape[0], _ptr(t), t.shape[0], 32, _ptr(idx),
                                                _ptr(dist), ctypes.byref(n)))
        return idx[: n.value], dist[: n.value]

    def bf_l2(self, query, train):
        """cv2.BFMatcher(cv2.NORM_L2).match on float32 descriptors: (train_idx, dist float32), bit-identical to cv2
        (integer-valued SIFT rows on the tensor cores, any other float data on the order-exact fp32 kernel)."""
        q, t = self._desc_l2(query), self._desc_l2(train, q_like=query)
        idx = np.empty(q.shape[0], np.int32)
        dist = np.empty(q.shape[0], np.float32)
        n = ctypes.c_int(0)
        self._check(self._lib.sfmgms_bf_l2(self._h, _ptr(q), q.shape[0], _ptr(t), t.shape[0], q.shape[1], _ptr(idx),
                                           _ptr(dist), ctypes.byref(n)))
        return idx[: n.value], dist[: n.value]

    def bf_hamming_crosscheck(self, query, train):
        q, t = self._desc(query), self._desc(train)
        idx = np.full(q.shape[0], -1, np.int32)
        dist = np.full(q.shape[0], -1, np.int32)
        keep = np.zeros(q.shape[0], np.uint8)
        self._check(self._lib.sfmgms_bf_hamming_crosscheck(self._h, _ptr(q), q.shape[0], _ptr(t), t.shape[0], 32,
                                                           _ptr(idx), _ptr(dist), _ptr(keep)))
        return idx, dist, keep.astype(bool)

    def bf_l2_crosscheck(self, query, train):
        """cv2.BFMatcher(cv2.NORM_L2, crossCheck=True) (FeatureMatchUtil.cpp:22-23): (train_idx, dist f32, keep)."""
        q, t = self._desc_l2(query), self._desc_l2(train, q_like=query)
        idx = np.full(q.shape[0], -1, np.int32)
        dist = np.full(q.shape[0], -1, np.float32)
        keep = np.zeros(q.shape[0], np.uint8)
        self._check(self._lib.sfmgms_bf_l2_crosscheck(self._h, _ptr(q), q.shape[0], _ptr(t), t.shape[0], q.shape[1], _ptr(idx),
                                                      _ptr(dist), _ptr(keep)))
        return idx, dist, keep.astype(bool)

    def brute_force_match(self, query, train, norm_type=NORM_L2, cross_check=True, distance_coef=4.0,
                          max_matching_size=500):
        """bruteForceMatch (FeatureMatchUtil.cpp:20-31): cross-checked NN, sort by distance, prune while
        front*coef < back, cap.  -> (query_idx, train_idx, dist float32), ascending distance (ties by queryIdx)."""
        if norm_type == NORM_L2:
            q, t = self._desc_l2(query), self._desc_l2(train, q_like=query)
            width = q.shape[1]
        else:
            q, t, width = self._desc(query), self._desc(train), 32
        cap = q.shape[0]
        qi = np.empty(cap, np.int32)
        ti = np.empty(cap, np.int32)
        d = np.empty(cap, np.float32)
        n = ctypes.c_int(0)
        self._check(self._lib.sfmgms_brute_force_match(self._h, int(norm_type), int(bool(cross_check)), _ptr(q), q.shape[0],
                                                       _ptr(t), t.shape[0], width, float(distance_coef),
                                                       int(max_matching_size), _ptr(qi), _ptr(ti), _ptr(d), cap,
                                                       ctypes.byref(n)))
        return qi[: n.value], ti[: n.value], d[: n.value]

    @staticmethod
    def _desc_l2(d, q_like=None):
        d = np.ascontiguousarray(d, dtype=np.float32)
        if d.ndim != 2 or d.shape[1] < 1:
            raise SfmGmsError(1, "L2 descriptors must be N x dim float32")
        if q_like is not None and np.ndim(q_like) == 2 and np.shape(q_like)[1] != d.shape[1]:
            raise SfmGmsError(1, "query and train descriptors differ in width")   # OpenCV asserts equal cols
        return d

    @staticmethod
    def _desc(d):
        d = np.ascontiguousarray(d)
        if d.dtype != np.uint8:
            raise SfmGmsError(1, "descriptors must be CV_8U (uint8), got %s" % d.dtype)  # OpenCV asserts type
        if d.ndim != 2 and d.size == 0:
            d = d.reshape(0, 32)
        if d.ndim != 2 or d.shape[1] != 32:
            raise SfmGmsError(1, "descriptors must be N x 32 bytes (256-bit), got %s" % (d.shape,))
        return d

    # -- (f4, first half) ORB descriptors on provided level-0 keypoints -----------------------------------
    def orb_compute(self, image, pts, angles=None, octaves=None):
        """cv2.ORB_create().compute(image, keypoints) for level-0 keypoints (DisparityUtil.cpp:107, 127-134).
        image: HxW (gray) or HxWx3 (BGR) uint8; pts (N, 2) float32 (x, y); angles (N,) degrees (None: all -1, the
        default KeyPoint angle).  -> (kept int32[n] = surviving input indices, desc uint8[n, 32])."""
        img = np.ascontiguousarray(image)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3):
            raise SfmGmsError(1, "image must be HxW or HxWx3 uint8")
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else 3
        pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
        n = pts.shape[0]
        rec = np.zeros((n, 4), np.float32)                    # x, y, angle, octave (int32 bits)
        rec[:, :2] = pts
        rec[:, 2] = -1.0 if angles is None else np.asarray(angles, np.float32).reshape(-1)
        if octaves is not None:
            rec[:, 3] = np.asarray(octaves, np.int32).reshape(-1).view(np.float32)
        kept = np.empty(n, np.int32)
        desc = np.empty((n, 32), np.uint8)
        nk = ctypes.c_int(0)
        self._check(self._lib.sfmgms_orb_compute(self._h, _ptr(img), w, h, ch, w * ch, _ptr(rec), n, 16, 8, 12, _ptr(kept),
                                                 _ptr(desc), ctypes.byref(nk)))
        return kept[: nk.value], desc[: nk.value]

    def orb_detect_and_compute(self, image, nfeatures=500, fast_threshold=20, with_descriptors=True, scale_factor=1.2,
                               nlevels=8, edge_threshold=31, score_type=0, first_level=0, wta_k=2, patch_size=31):
        """cv2.ORB_create(nfeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize,
        fastThreshold).detectAndCompute(image, None) (DisparityUtil.cpp:107, 139-140 use the defaults).
        -> (kp float32[n, 6] = x, y, size, angle, response, octave in OpenCV's output order, desc uint8[n, 32] | None)"""
        img = np.ascontiguousarray(image)
        if img.dtype != np.uint8 or img.ndim not in (2, 3) or (img.ndim == 3 and img.shape[2] != 3):
            raise SfmGmsError(1, "image must be HxW or HxWx3 uint8")
        h, w = img.shape[:2]
        ch = 1 if img.ndim == 2 else 3
        prm = np.zeros(9, np.int32)
        prm[:] = [int(nfeatures), 0, int(nlevels), int(edge_threshold), int(first_level), int(wta_k), int(score_type),
                  int(patch_size), int(fast_threshold)]
        prm[1:2] = np.array([scale_factor], np.float32).view(np.int32)
        cap = max(2 * int(nfeatures), 64)
        for _ in range(2):
            rec = np.zeros((cap, 7), np.float32)              # cv::KeyPoint records (28 bytes)
            desc = np.zeros((cap, 32), np.uint8) if with_descriptors else None
            n = ctypes.c_int(0)
            rc = self._lib.sfmgms_orb_detect_and_compute_ex(self._h, _ptr(img), w, h, ch, w * ch, _ptr(prm), _ptr(rec),
                                                            _ptr(desc) if with_descriptors else None, cap, ctypes.byref(n))
            if rc == 1 and n.value > cap:                      # ties at a level's cut: retry with the reported size
                cap = n.value
                continue
            self._check(rc)
            break
        kp = np.empty((n.value, 6), np.float32)
        kp[:, :5] = rec[: n.value, :5]
        kp[:, 5] = rec[: n.value, 5].view(np.int32)
        return kp, (desc[: n.value] if with_descriptors else None)

    # -- stage 2 ------------------------------------------------------------------------------------
    def gms(self, size1, size2, kp1, kp2, query_idx, train_idx, with_rotation=False, with_scale=False,
            threshold_factor=6.0):
        """-> dict(mask bool[mask_len], n_inliers, best_hyp).  OpenCV argument order (rotation, scale)."""
        w1, h1 = _size(size1)
        w2, h2 = _size(size2)
        k1, k2 = _kp_xy(kp1), _kp_xy(kp2)
        qi = np.ascontiguousarray(query_idx, dtype=np.int32)
        ti = np.ascontiguousarray(train_idx, dtype=np.int32)
        n = qi.shape[0]
        mask = np.zeros(max(n, 1), np.uint8)
        ml, ni, bh = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(-1)
        self._check(self._lib.sfmgms_gms(self._h, w1, h1, w2, h2, _ptr(k1), k1.shape[0], 8, _ptr(k2), k2.shape[0], 8,
                                         _ptr(qi), _ptr(ti), 4, n, int(bool(with_rotation)), int(bool(with_scale)),
                                         float(threshold_factor), _ptr(mask), ctypes.byref(ml), ctypes.byref(ni),
                                         ctypes.byref(bh)))
        return dict(mask=mask[: ml.value].astype(bool), n_inliers=ni.value, best_hyp=bh.value)

    def gms_hypotheses(self, size1, size2, kp1, kp2, query_idx, train_idx, threshold_factor=6.0):
        """-> int32[40]: inlier count of every (scale, rotation) hypothesis, scale-major (GMSMatcher::run per hypothesis)."""
        w1, h1 = _size(size1)
        w2, h2 = _size(size2)
        k1, k2 = _kp_xy(kp1), _kp_xy(kp2)
        qi = np.ascontiguousarray(query_idx, dtype=np.int32)
        ti = np.ascontiguousarray(train_idx, dtype=np.int32)
        counts = np.zeros(40, np.int32)
        self._check(self._lib.sfmgms_gms_hypotheses(self._h, w1, h1, w2, h2, _ptr(k1), k1.shape[0], 8, _ptr(k2), k2.shape[0], 8,
                                                    _ptr(qi), _ptr(ti), 4, qi.shape[0], float(threshold_factor), _ptr(counts)))
        return counts

    def kernel_times(self):
        """OPT_TIMING = 2: {kernel name: (total ms, launches)} since the last call."""
        buf = ctypes.create_string_buffer(8192)
        self._check(self._lib.sfmgms_kernel_times(self._h, buf, 8192))
        out = {}
        for item in buf.value.decode().split(";"):
            if item:
                name, ms, n = item.split(":")
                out[name] = (float(ms), int(n))
        return out

    @property
    def device_bytes(self):
        return int(self._lib.sfmgms_device_bytes(self._h))

    # -- fused pair ---------------------------------------------------------------------------------
    def match_pair(self, desc1, desc2, kp1, kp2, size1, size2, with_rotation=False, with_scale=False,
                   threshold_factor=6.0):
        d1, d2 = self._desc(desc1), self._desc(desc2)
        k1, k2 = _kp_xy(kp1), _kp_xy(kp2)
        if k1.shape[0] != d1.shape[0] or k2.shape[0] != d2.shape[0]:
            raise SfmGmsError(1, "keypoint / descriptor row counts differ")
        w1, h1 = _size(size1)
        w2, h2 = _size(size2)
        n1 = d1.shape[0]
        nm = n1 if d2.shape[0] else 0
        idx = np.empty(max(n1, 1), np.int32)
        dist = np.empty(max(n1, 1), np.int32)
        mask = np.zeros(max(n1, 1), np.uint8)
        ml, ni, bh = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(-1)
        self._check(self._lib.sfmgms_match_pair(self._h, _ptr(d1), n1, _ptr(d2), d2.shape[0], 32, _ptr(k1), 8,
                                                _ptr(k2), 8, w1, h1, w2, h2, int(bool(with_rotation)),
                                                int(bool(with_scale)), float(threshold_factor), _ptr(idx),
                                                _ptr(dist), _ptr(mask), ctypes.byref(ml), ctypes.byref(ni),
                                                ctypes.byref(bh)))
        return dict(train_idx=idx[:nm], dist=dist[:nm], mask=mask[: ml.value].astype(bool), n_inliers=ni.value,
                    best_hyp=bh.value)

    # -- multi-pair ---------------------------------------------------------------------------------
    def set_images(self, kp_offsets, desc, kp_xy, sizes_wh):
        """Register an image set from HOST numpy arrays (copied to the device)."""
        off = np.ascontiguousarray(kp_offsets, dtype=np.int64)
        d = self._desc(desc)
        k = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
        s = np.ascontiguousarray(sizes_wh, dtype=np.int32).reshape(-1, 2)
        n_images = off.shape[0] - 1
        if s.shape[0] != n_images or d.shape[0] != off[-1] or k.shape[0] != off[-1]:
            raise SfmGmsError(1, "image-set array shapes are inconsistent")
        self._check(self._lib.sfmgms_set_images(self._h, n_images, _ptr(off), _ptr(d), _ptr(k), _ptr(s), SFMGMS_HOST))
        self._offsets = off
        self._keep = None

    def set_images_device(self, kp_offsets, desc_ptr, kp_ptr, sizes_wh, keepalive=None):
        """Adopt device buffers (e.g. torch tensors that received an NCCL broadcast) without a copy."""
        off = np.ascontiguousarray(kp_offsets, dtype=np.int64)
        s = np.ascontiguousarray(sizes_wh, dtype=np.int32).reshape(-1, 2)
        self._check(self._lib.sfmgms_set_images(self._h, off.shape[0] - 1, _ptr(off), ctypes.c_void_p(int(desc_ptr)),
                                                ctypes.c_void_p(int(kp_ptr)), _ptr(s), SFMGMS_DEVICE))
        self._offsets = off
        self._keep = keepalive

    def match_offsets(self, pairs):
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        out = np.zeros(pr.shape[0] + 1, np.int64)
        self._check(self._lib.sfmgms_match_offsets(self._h, _ptr(pr), pr.shape[0], _ptr(out)))
        return out

    def match_pairs(self, pairs, with_rotation=False, with_scale=False, threshold_factor=6.0, want_matches=True,
                    want_mask=True):
        """Host outputs.  -> dict(n_inliers, best_hyp, mask_len, offsets[, train_idx, dist][, mask])."""
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pr.shape[0]
        off = self.match_offsets(pr)
        total = int(off[-1])
        ninl = np.zeros(n, np.int32)
        bh = np.zeros(n, np.int32)
        ml = np.zeros(n, np.int32)
        ti = np.empty(total, np.int32) if want_matches else None
        di = np.empty(total, np.int32) if want_matches else None
        mk = np.zeros(total, np.uint8) if want_mask else None
        self._check(self._lib.sfmgms_match_pairs(self._h, _ptr(pr), n, int(bool(with_rotation)), int(bool(with_scale)),
                                                 float(threshold_factor), SFMGMS_HOST, _ptr(ninl), _ptr(bh), _ptr(ml),
                                                 _ptr(ti), _ptr(di), _ptr(mk)))
        out = dict(n_inliers=ninl, best_hyp=bh, mask_len=ml, offsets=off)
        if want_matches:
            out["train_idx"], out["dist"] = ti, di
        if want_mask:
            out["mask"] = mk
        return out

    def match_pairs_compact(self, pairs, with_rotation=False, with_scale=False, threshold_factor=6.0, capacity=None,
                            want_matches=True, want_points=True, index_pairs=False):
        """Host outputs: every pair's matchesGMS (cv::DMatch records) and inlier coordinates, back to back.
        -> dict(n_inliers, best_hyp, offsets int64[n+1], n_total[, matches (DMATCH_DT)][, pts1, pts2 float32[n,2]]).
        index_pairs: `matches` holds int32 (queryIdx, trainIdx) rows instead (SFMGMS_OPT_COMPACT_RECORD = 1)."""
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pr.shape[0]
        if capacity is None:
            capacity = int(self.match_offsets(pr)[-1])
        ninl = np.zeros(n, np.int32)
        bh = np.zeros(n, np.int32)
        off = np.zeros(n + 1, np.int64)
        m = (np.zeros((max(capacity, 1), 2), np.int32) if index_pairs else np.zeros(max(capacity, 1), DMATCH_DT)) if want_matches else None
        p1 = np.zeros((max(capacity, 1), 2), np.float32) if want_points else None
        p2 = np.zeros((max(capacity, 1), 2), np.float32) if want_points else None
        tot = ctypes.c_int64(0)
        self.set_option(OPT_COMPACT_RECORD, 1 if index_pairs else 0)
        try:
            self._check(self._lib.sfmgms_match_pairs_compact(self._h, _ptr(pr), n, int(bool(with_rotation)), int(bool(with_scale)),
                                                             float(threshold_factor), SFMGMS_HOST, _ptr(ninl), _ptr(bh), _ptr(off),
                                                             _ptr(m), _ptr(p1), _ptr(p2), int(capacity), ctypes.byref(tot)))
        finally:
            self.set_option(OPT_COMPACT_RECORD, 0)
        out = dict(n_inliers=ninl, best_hyp=bh, offsets=off, n_total=tot.value)
        if want_matches:
            out["matches"] = m[: tot.value]
        if want_points:
            out["pts1"], out["pts2"] = p1[: tot.value], p2[: tot.value]
        return out

    def match_pairs_compact_raw(self, pairs_np, with_rotation, with_scale, threshold_factor, out_location, capacity,
                                n_inliers=0, best_hyp=0, offsets=0, matches=0, pts1=0, pts2=0):
        """Raw-pointer form (ints are addresses; 0 = NULL).  -> n_total"""
        p = lambda v: ctypes.c_void_p(int(v)) if v else None  # noqa: E731
        tot = ctypes.c_int64(0)
        self._check(self._lib.sfmgms_match_pairs_compact(self._h, _ptr(pairs_np), pairs_np.shape[0], int(with_rotation),
                                                         int(with_scale), float(threshold_factor), int(out_location),
                                                         p(n_inliers), p(best_hyp), p(offsets), p(matches), p(pts1), p(pts2),
                                                         int(capacity), ctypes.byref(tot)))
        return tot.value

    def match_pairs_compact_async_raw(self, pairs_np, with_rotation, with_scale, threshold_factor, capacity, n_inliers=0, best_hyp=0,
                                      offsets=0, matches=0, pts1=0, pts2=0):
        """Device pointers only; returns at once.  Finish with wait()."""
        p = lambda v: ctypes.c_void_p(int(v)) if v else None  # noqa: E731
        self._check(self._lib.sfmgms_match_pairs_compact_async(self._h, _ptr(pairs_np), pairs_np.shape[0], int(with_rotation),
                                                               int(with_scale), float(threshold_factor), p(n_inliers), p(best_hyp),
                                                               p(offsets), p(matches), p(pts1), p(pts2), int(capacity)))

    def match_pairs_async_raw(self, pairs_np, with_rotation, with_scale, threshold_factor, n_inliers=0, best_hyp=0, mask_len=0,
                              train_idx=0, dist=0, mask=0):
        p = lambda v: ctypes.c_void_p(int(v)) if v else None  # noqa: E731
        self._check(self._lib.sfmgms_match_pairs_async(self._h, _ptr(pairs_np), pairs_np.shape[0], int(with_rotation), int(with_scale),
                                                       float(threshold_factor), p(n_inliers), p(best_hyp), p(mask_len), p(train_idx),
                                                       p(dist), p(mask)))

    def wait(self):
        """Completes an asynchronous pair list -> total number of inliers."""
        tot = ctypes.c_int64(0)
        self._check(self._lib.sfmgms_wait(self._h, ctypes.byref(tot)))
        return tot.value

    def match_pairs_raw(self, pairs_np, with_rotation, with_scale, threshold_factor, out_location, n_inliers=0,
                        best_hyp=0, mask_len=0, train_idx=0, dist=0, mask=0):
        """Raw-pointer form (ints are addresses; 0 = NULL) for callers that manage their own buffers
        (bench.py: pinned host tensors or device tensors)."""
        p = lambda v: ctypes.c_void_p(int(v)) if v else None  # noqa: E731
        self._check(self._lib.sfmgms_match_pairs(self._h, _ptr(pairs_np), pairs_np.shape[0], int(with_rotation),
                                                 int(with_scale), float(threshold_factor), int(out_location),
                                                 p(n_inliers), p(best_hyp), p(mask_len), p(train_idx), p(dist),
                                                 p(mask)))

    def set_images_raw(self, kp_offsets, desc_ptr, kp_ptr, sizes_wh, location, keepalive=None):
        off = np.ascontiguousarray(kp_offsets, dtype=np.int64)
        s = np.ascontiguousarray(sizes_wh, dtype=np.int32).reshape(-1, 2)
        self._check(self._lib.sfmgms_set_images(self._h, off.shape[0] - 1, _ptr(off), ctypes.c_void_p(int(desc_ptr)),
                                                ctypes.c_void_p(int(kp_ptr)), _ptr(s), int(location)))
        self._offsets = off
        self._keep = keepalive

    def set_images_from_pixels(self, images, nfeatures=500, fast_threshold=20, scale_factor=1.2, nlevels=8,
                               edge_threshold=31, score_type=0, wta_k=2, patch_size=31):
        """ORB on every image (HxW or HxWx3 uint8 arrays) -> the image set of match_pairs.  -> kp_offsets int64[n+1]"""
        imgs = [np.ascontiguousarray(im) for im in images]
        for im in imgs:
            if im.dtype != np.uint8 or im.ndim not in (2, 3) or (im.ndim == 3 and im.shape[2] != 3):
                raise SfmGmsError(1, "images must be HxW or HxWx3 uint8")
        n = len(imgs)
        ptrs = (ctypes.c_void_p * max(n, 1))(*[im.ctypes.data for im in imgs])
        w = np.array([im.shape[1] for im in imgs], np.int32)
        h = np.array([im.shape[0] for im in imgs], np.int32)
        ch = np.array([1 if im.ndim == 2 else 3 for im in imgs], np.int32)
        prm = np.zeros(9, np.int32)
        prm[:] = [int(nfeatures), 0, int(nlevels), int(edge_threshold), 0, int(wta_k), int(score_type), int(patch_size),
                  int(fast_threshold)]
        prm[1:2] = np.array([scale_factor], np.float32).view(np.int32)
        off = np.zeros(n + 1, np.int64)
        self._check(self._lib.sfmgms_set_images_from_pixels(self._h, n, ctypes.cast(ptrs, ctypes.c_void_p), _ptr(w), _ptr(h), _ptr(ch),
                                                            None, _ptr(prm), _ptr(off)))
        self._offsets = off
        self._keep = None
        return off

    def get_image_keypoints(self, image):
        """-> float32[n, 6] = x, y, size, angle, response, octave of one image of the last set_images_from_pixels"""
        n = ctypes.c_int(0)
        self._lib.sfmgms_get_image_keypoints(self._h, int(image), None, 0, ctypes.byref(n))
        rec = np.zeros((max(n.value, 1), 7), np.float32)
        self._check(self._lib.sfmgms_get_image_keypoints(self._h, int(image), _ptr(rec), max(n.value, 1), ctypes.byref(n)))
        kp = np.empty((n.value, 6), np.float32)
        kp[:, :5] = rec[: n.value, :5]
        kp[:, 5] = rec[: n.value, 5].view(np.int32)
        return kp

    def match_image_set(self, kp_offsets, desc, kp_xy, sizes_wh, pairs, with_rotation=False, with_scale=False,
                        threshold_factor=6.0):
        """set_images + match_pairs in one internally pipelined call (H2D / compute / D2H overlap)."""
        off = np.ascontiguousarray(kp_offsets, dtype=np.int64)
        d = self._desc(desc)
        k = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
        s = np.ascontiguousarray(sizes_wh, dtype=np.int32).reshape(-1, 2)
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pr.shape[0]
        if s.shape[0] != off.shape[0] - 1 or d.shape[0] != off[-1] or k.shape[0] != off[-1]:
            raise SfmGmsError(1, "image-set array shapes are inconsistent")
        if n and (pr.min() < 0 or pr.max() >= s.shape[0]):
            raise SfmGmsError(1, "a pair references an image outside the set")
        mo = np.zeros(n + 1, np.int64)
        np.cumsum((off[1:] - off[:-1])[pr[:, 0]], out=mo[1:])
        total = int(mo[-1])
        out = dict(n_inliers=np.zeros(n, np.int32), best_hyp=np.zeros(n, np.int32), mask_len=np.zeros(n, np.int32),
                   offsets=mo, train_idx=np.empty(total, np.int32), dist=np.empty(total, np.int32),
                   mask=np.zeros(total, np.uint8))
        self._check(self._lib.sfmgms_match_image_set(self._h, s.shape[0], _ptr(off), _ptr(d), _ptr(k), _ptr(s), _ptr(pr), n,
                                                     int(bool(with_rotation)), int(bool(with_scale)),
                                                     float(threshold_factor), _ptr(out["n_inliers"]),
                                                     _ptr(out["best_hyp"]), _ptr(out["mask_len"]), _ptr(out["train_idx"]),
                                                     _ptr(out["dist"]), _ptr(out["mask"])))
        self._offsets = off
        return out

    def match_image_set_raw(self, kp_offsets, desc_ptr, kp_ptr, sizes_wh, pairs_np, with_rotation, with_scale,
                            threshold_factor, n_inliers=0, best_hyp=0, mask_len=0, train_idx=0, dist=0, mask=0):
        """Raw-pointer form (host addresses; 0 = NULL) for callers with their own pinned buffers (bench.py)."""
        p = lambda v: ctypes.c_void_p(int(v)) if v else None  # noqa: E731
        off = np.ascontiguousarray(kp_offsets, dtype=np.int64)
        s = np.ascontiguousarray(sizes_wh, dtype=np.int32).reshape(-1, 2)
        self._check(self._lib.sfmgms_match_image_set(self._h, off.shape[0] - 1, _ptr(off), p(desc_ptr), p(kp_ptr), _ptr(s),
                                                     _ptr(pairs_np), pairs_np.shape[0], int(with_rotation),
                                                     int(with_scale), float(threshold_factor), p(n_inliers), p(best_hyp),
                                                     p(mask_len), p(train_idx), p(dist), p(mask)))
        self._offsets = off

    def inlier_points(self, pair_index, capacity):
        p1 = np.empty((max(capacity, 1), 2), np.float32)
        p2 = np.empty((max(capacity, 1), 2), np.float32)
        n = ctypes.c_int(0)
        self._check(self._lib.sfmgms_inlier_points(self._h, int(pair_index), _ptr(p1), _ptr(p2), int(capacity),
                                                   ctypes.byref(n)))
        m = min(n.value, capacity)
        return p1[:m], p2[:m], n.value


_default = threading.local()


class MultiContext:
    """One process, several GPUs (sfmgms_multi_*): the set is broadcast once, the pair list is sharded."""

    def __init__(self, devices=None, n_devices=0):
        self._lib = load_library()
        h = ctypes.c_void_p()
        dv = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        rc = self._lib.sfmgms_multi_create(ctypes.byref(h), _ptr(dv), len(dv) if dv is not None else int(n_devices))
        if rc:
            raise SfmGmsError(rc, self._lib.sfmgms_multi_last_error(None).decode())
        self._h = h
        self.n_devices = self._lib.sfmgms_multi_device_count(h)
        self._offsets = None

    def close(self):
        if self._h:
            self._lib.sfmgms_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise SfmGmsError(rc, self._lib.sfmgms_multi_last_error(self._h).decode())

    def set_option(self, key, value):
        for i in range(self.n_devices):
            c = ctypes.c_void_p(self._lib.sfmgms_multi_context(self._h, i))
            rc = self._lib.sfmgms_set_option(c, int(key), int(value))
            if rc:
                raise SfmGmsError(rc, self._lib.sfmgms_last_error(c).decode())

    @property
    def last_broadcast_ms(self):
        return float(self._lib.sfmgms_multi_last_broadcast_ms(self._h))

    def set_images(self, kp_offsets, desc, kp_xy, sizes_wh):
        off = np.ascontiguousarray(kp_offsets, dtype=np.int64)
        d = Context._desc(desc)
        k = np.ascontiguousarray(kp_xy, dtype=np.float32).reshape(-1, 2)
        s = np.ascontiguousarray(sizes_wh, dtype=np.int32).reshape(-1, 2)
        self._check(self._lib.sfmgms_multi_set_images(self._h, off.shape[0] - 1, _ptr(off), _ptr(d), _ptr(k), _ptr(s)))
        self._offsets = off

    def _need_set(self):
        if self._offsets is None:
            raise SfmGmsError(6, "set_images has not been called")

    def match_pairs(self, pairs, with_rotation=False, with_scale=False, threshold_factor=6.0):
        self._need_set()
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pr.shape[0]
        rows = (self._offsets[pr[:, 0] + 1] - self._offsets[pr[:, 0]]) if n else np.zeros(0, np.int64)
        off = np.concatenate([[0], np.cumsum(rows)]).astype(np.int64)
        total = int(off[-1])
        ninl, bh, ml = (np.zeros(n, np.int32) for _ in range(3))
        ti, di = np.empty(total, np.int32), np.empty(total, np.int32)
        mk = np.zeros(total, np.uint8)
        self._check(self._lib.sfmgms_multi_match_pairs(self._h, _ptr(pr), n, int(bool(with_rotation)), int(bool(with_scale)),
                                                       float(threshold_factor), _ptr(ninl), _ptr(bh), _ptr(ml), _ptr(ti), _ptr(di),
                                                       _ptr(mk)))
        return dict(n_inliers=ninl, best_hyp=bh, mask_len=ml, offsets=off, train_idx=ti, dist=di, mask=mk)

    def match_pairs_compact(self, pairs, with_rotation=False, with_scale=False, threshold_factor=6.0, capacity=None):
        """-> dict(n_inliers, best_hyp, begin int64[n], n_total, matches (DMATCH_DT), pts1, pts2); pair p owns rows
        [begin[p], begin[p] + n_inliers[p])."""
        self._need_set()
        pr = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 2)
        n = pr.shape[0]
        if capacity is None:
            capacity = int((self._offsets[pr[:, 0] + 1] - self._offsets[pr[:, 0]]).sum()) if n else 0
        ninl, bh = np.zeros(n, np.int32), np.zeros(n, np.int32)
        beg = np.zeros(n, np.int64)
        m = np.zeros(max(capacity, 1), DMATCH_DT)
        p1, p2 = np.zeros((max(capacity, 1), 2), np.float32), np.zeros((max(capacity, 1), 2), np.float32)
        tot = ctypes.c_int64(0)
        self._check(self._lib.sfmgms_multi_match_pairs_compact(self._h, _ptr(pr), n, int(bool(with_rotation)), int(bool(with_scale)),
                                                               float(threshold_factor), _ptr(ninl), _ptr(bh), _ptr(beg), _ptr(m),
                                                               _ptr(p1), _ptr(p2), int(capacity), ctypes.byref(tot)))
        return dict(n_inliers=ninl, best_hyp=bh, begin=beg, n_total=tot.value, matches=m[: tot.value], pts1=p1[: tot.value],
                    pts2=p2[: tot.value])


def default_context(device=0):
    ctxs = getattr(_default, "ctxs", None)
    if ctxs is None:
        ctxs = _default.ctxs = {}
    if device not in ctxs:
        ctxs[device] = Context(device)
    return ctxs[device]


class BFMatcher:
    """cv::BFMatcher look-alike for NORM_HAMMING (FeatureMatchUtil.cpp:22, 66)."""

    def __init__(self, normType=NORM_HAMMING, crossCheck=False, ctx=None):
        if normType not in (NORM_HAMMING, NORM_L2):
            raise SfmGmsError(1, "implemented: NORM_HAMMING and NORM_L2 (SIFT), with/without crossCheck; "
                                 "got normType=%r" % (normType,))
        self.normType = normType
        self.crossCheck = bool(crossCheck)
        self._ctx = ctx

    @staticmethod
    def create(normType=NORM_HAMMING, crossCheck=False):
        return BFMatcher(normType, crossCheck)

    def match(self, queryDescriptors, trainDescriptors):
        """-> list[DMatch] exactly as cv2 returns it (query order; cross-check drops non-mutual rows)."""
        ctx = self._ctx or default_context()
        if self.normType == NORM_L2 and self.crossCheck:
            idx, dist, keep = ctx.bf_l2_crosscheck(queryDescriptors, trainDescriptors)
            return [DMatch(int(i), int(idx[i]), 0, float(dist[i])) for i in np.nonzero(keep)[0]]
        if self.normType == NORM_L2:
            idx, dist = ctx.bf_l2(queryDescriptors, trainDescriptors)
            return [DMatch(i, int(idx[i]), 0, float(dist[i])) for i in range(len(idx))]
        if self.crossCheck:
            idx, dist, keep = ctx.bf_hamming_crosscheck(queryDescriptors, trainDescriptors)
            return [DMatch(int(i), int(idx[i]), 0, float(dist[i])) for i in np.nonzero(keep)[0]]
        idx, dist = ctx.bf_hamming(queryDescriptors, trainDescriptors)
        return [DMatch(i, int(idx[i]), 0, float(dist[i])) for i in range(len(idx))]


class KeyPoint:
    """cv2.KeyPoint-shaped record returned by ORB.detect / detectAndCompute."""
    __slots__ = ("pt", "size", "angle", "response", "octave", "class_id")

    def __init__(self, x, y, size, angle=-1.0, response=0.0, octave=0, class_id=-1):
        self.pt, self.size, self.angle, self.response, self.octave, self.class_id = (x, y), size, angle, response, octave, class_id

    def __repr__(self):
        return "KeyPoint(pt=%r, size=%r, angle=%r, response=%r, octave=%r)" % (self.pt, self.size, self.angle, self.response, self.octave)


class ORB:
    """cv::ORB look-alike: ``ORB_create(nfeatures)`` with the other defaults of ``ORB::create()`` (DisparityUtil.cpp:107:
    scaleFactor 1.2, 8 levels, edgeThreshold 31, HARRIS_SCORE, patchSize 31), ``setFastThreshold``, ``detect``,
    ``compute`` and ``detectAndCompute`` -- bit-identical to cv2 (keypoint order included)."""

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures=500, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2, scoreType=0,
                 patchSize=31, fastThreshold=20, ctx=None):
        self._ctx = ctx
        self._nfeatures = int(nfeatures)
        self._fast = int(fastThreshold)
        self._kw = dict(scale_factor=scaleFactor, nlevels=nlevels, edge_threshold=edgeThreshold, score_type=scoreType,
                        first_level=firstLevel, wta_k=WTA_K, patch_size=patchSize)

    @staticmethod
    def create(nfeatures=500, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2, scoreType=0, patchSize=31,
               fastThreshold=20):
        return ORB(nfeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize, fastThreshold)

    def setFastThreshold(self, t):
        self._fast = int(t)

    def getFastThreshold(self):
        return self._fast

    def descriptorSize(self):
        return 32

    def detectAndCompute(self, image, mask=None):
        if mask is not None:
            raise SfmGmsError(1, "masks are not implemented")
        ctx = self._ctx or default_context()
        kp, desc = ctx.orb_detect_and_compute(image, self._nfeatures, self._fast, True, **self._kw)
        return [KeyPoint(float(r[0]), float(r[1]), float(r[2]), float(r[3]), float(r[4]), int(r[5])) for r in kp], desc

    def detect(self, image, mask=None):
        if mask is not None:
            raise SfmGmsError(1, "masks are not implemented")
        ctx = self._ctx or default_context()
        kp, _ = ctx.orb_detect_and_compute(image, self._nfeatures, self._fast, False, **self._kw)
        return [KeyPoint(float(r[0]), float(r[1]), float(r[2]), float(r[3]), float(r[4]), int(r[5])) for r in kp]

    def compute(self, image, keypoints):
        """-> (kept keypoints, descriptors) as cv2 returns them.  keypoints: list of cv2.KeyPoint-like objects
        (.pt, .angle, .octave) -> a list comes back; or a tuple (pts[, angles[, octaves]]) of arrays -> the kept indices."""
        ctx = self._ctx or default_context()
        if isinstance(keypoints, tuple):
            pts = keypoints[0]
            ang = keypoints[1] if len(keypoints) > 1 else None
            octv = keypoints[2] if len(keypoints) > 2 else None
            return ctx.orb_compute(image, pts, ang, octv)
        pts = np.array([k.pt for k in keypoints], np.float32).reshape(-1, 2)
        ang = np.array([k.angle for k in keypoints], np.float32)
        octv = np.array([getattr(k, "octave", 0) for k in keypoints], np.int32)
        kept, desc = ctx.orb_compute(image, pts, ang, octv)
        return [keypoints[i] for i in kept], desc


def ORB_create(nfeatures=500, scaleFactor=1.2, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2, scoreType=0, patchSize=31,
               fastThreshold=20, ctx=None):
    return ORB(nfeatures, scaleFactor, nlevels, edgeThreshold, firstLevel, WTA_K, scoreType, patchSize, fastThreshold, ctx)


def bruteForceMatch(desc1, desc2, ctx=None, kDistanceCoef=4.0, kMaxMatchingSize=500):
    """The reference's bruteForceMatch(desc1, desc2, matches) (FeatureMatchUtil.cpp:20-31; constants from
    FeatureMatchUtil.h:17-18): returns the list of DMatch it leaves in ``matches``.  float32 N x 128 descriptors
    take the reference's NORM_L2 path; uint8 N x 32 descriptors the Hamming one."""
    ctx = ctx or default_context()
    norm = NORM_HAMMING if np.asarray(desc1).dtype == np.uint8 else NORM_L2
    q, t, d = ctx.brute_force_match(desc1, desc2, norm, True, kDistanceCoef, kMaxMatchingSize)
    return [DMatch(int(q[i]), int(t[i]), 0, float(d[i])) for i in range(len(q))]


def match(desc1, desc2, kDistanceCoef, kMaxMatchingSize, ctx=None):
    """The reference's inline match(desc1, desc2, matches, kDistanceCoef, kMaxMatchingSize)
    (FeatureMatchUtil.cpp:38-50): BFMatcher::create() = NORM_L2 without cross-check, then the same sort/prune/cap."""
    ctx = ctx or default_context()
    norm = NORM_HAMMING if np.asarray(desc1).dtype == np.uint8 else NORM_L2
    q, t, d = ctx.brute_force_match(desc1, desc2, norm, False, kDistanceCoef, kMaxMatchingSize)
    return [DMatch(int(q[i]), int(t[i]), 0, float(d[i])) for i in range(len(q))]


def matchGMS(size1, size2, keypoints1, keypoints2, matches1to2, withRotation=False, withScale=False,
             thresholdFactor=6.0, ctx=None):
    """cv::xfeatures2d::matchGMS: returns the sub-list of matches1to2 kept by GMS, input order preserved."""
    ctx = ctx or default_context()
    q, t = _match_idx(matches1to2)
    r = ctx.gms(size1, size2, keypoints1, keypoints2, q, t, withRotation, withScale, thresholdFactor)
    keep = np.nonzero(r["mask"])[0]
    if isinstance(matches1to2, tuple):
        return q[keep], t[keep]
    return [matches1to2[i] for i in keep]


class gms_matcher:
    """Upstream header-only API: gms_matcher(vkp1, size1, vkp2, size2, vDMatches).GetInlierMask(...)."""

    def __init__(self, vkp1, size1, vkp2, size2, vDMatches, ctx=None):
        self._args = (size1, size2, _kp_xy(vkp1), _kp_xy(vkp2)) + _match_idx(vDMatches)
        self._ctx = ctx

    def GetInlierMask(self, WithScale=False, WithRotation=False):
        """-> (num_inliers, vbInliers).  NOTE upstream order: scale first, rotation second; threshold 6."""
        ctx = self._ctx or default_context()
        s1, s2, k1, k2, q, t = self._args
        r = ctx.gms(s1, s2, k1, k2, q, t, with_rotation=WithRotation, with_scale=WithScale, threshold_factor=6.0)
        return r["n_inliers"], r["mask"]
