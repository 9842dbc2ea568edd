"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (ctypes), against
the CPU oracle on the same inputs and against the committed golden fixtures.  Bit-exact: the path is
integer (indices, distances, masks); the three FP decisions inside GMS are reproduced exactly.
"""
import numpy as np
import pytest

from conftest import load_golden
from microcases import MICROCASES, run_microcase

pytestmark = pytest.mark.gpu

KERNELS = ["popc", "tc", "fp4"]


def _select(ctx, kernel):
    from sfm_gms_b200 import api

    try:
        ctx.set_option(api.OPT_HAMMING_KERNEL, {"popc": api.HAMMING_POPC, "tc": api.HAMMING_TC, "fp4": api.HAMMING_FP4}[kernel])
    except api.SfmGmsError as e:
        pytest.skip("kernel %s not available: %s" % (kernel, e))


@pytest.fixture(params=KERNELS)
def kctx(ctx, request):
    from sfm_gms_b200 import api

    _select(ctx, request.param)
    yield ctx
    ctx.set_option(api.OPT_HAMMING_KERNEL, api.HAMMING_AUTO)


# ---- stage 1 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["pikabun12", "disparityLR", "view01_2k", "bun12_rot180_3k"])
def test_bf_golden_cv2(kctx, name):
    g = load_golden(name)
    idx, dist = kctx.bf_hamming(g["desc1"], g["desc2"])
    assert np.array_equal(idx, g["bf_train"]) and np.array_equal(dist, g["bf_dist"])


def test_bf_ties_golden_cv2(kctx):
    g = load_golden("bf_ties")
    idx, dist = kctx.bf_hamming(g["q"], g["t"])
    assert np.array_equal(idx, g["train"]) and np.array_equal(dist, g["dist"])
    idx, dist = kctx.bf_hamming(g["q2"], g["t2"])
    assert np.array_equal(idx, g["train2"]) and np.array_equal(dist, g["dist2"])


@pytest.mark.parametrize("nq,nt", [(1, 1), (1, 5000), (5000, 1), (31, 33), (127, 129), (128, 128), (129, 255),
                                   (512, 256), (513, 257), (1000, 999), (4097, 2049), (10000, 10000)])
def test_bf_ragged_vs_oracle(kctx, oracle_mod, nq, nt):
    rng = np.random.default_rng(nq * 7919 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    if nt > 8:  # plant exact and near duplicates to exercise tie-breaks across tile borders
        t[nt // 2] = t[1]
        t[nt - 1] = t[0]
        q[0] = t[0]
        q[nq // 2] = t[1]
    idx, dist = kctx.bf_hamming(q, t)
    oi, od = oracle_mod.bf_hamming(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


def test_bf_low_entropy_many_ties(kctx, oracle_mod):
    rng = np.random.default_rng(3)
    q = (rng.integers(0, 2, (3000, 32), dtype=np.uint8) * 255).astype(np.uint8)
    t = (rng.integers(0, 2, (2500, 32), dtype=np.uint8) * 255).astype(np.uint8)
    idx, dist = kctx.bf_hamming(q, t)
    oi, od = oracle_mod.bf_hamming(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    z = np.zeros((700, 32), np.uint8)
    f = np.full((900, 32), 255, np.uint8)
    idx, dist = kctx.bf_hamming(z, f)
    assert (idx == 0).all() and (dist == 256).all()   # maximum distance, all tied: index 0


def test_bf_empty_and_limits(kctx):
    from sfm_gms_b200 import SfmGmsError

    q = np.zeros((5, 32), np.uint8)
    idx, dist = kctx.bf_hamming(q, np.zeros((0, 32), np.uint8))
    assert len(idx) == 0 and len(dist) == 0
    idx, dist = kctx.bf_hamming(np.zeros((0, 32), np.uint8), q)
    assert len(idx) == 0
    with pytest.raises(SfmGmsError) as e:
        kctx.bf_hamming(q, np.zeros((1 << 18, 32), np.uint8))
    assert e.value.code == 2
    with pytest.raises(SfmGmsError):
        kctx.bf_hamming(np.zeros((5, 16), np.uint8), q)


def test_bf_max_train_rows(kctx, oracle_mod):
    """train rows = 2^18 - 1 (OpenCV's maximum); sampled queries against the oracle."""
    rng = np.random.default_rng(99)
    nt = (1 << 18) - 1
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    q[7] = t[nt - 1]
    t[5] = t[nt - 2]
    q[8] = t[nt - 2]
    idx, dist = kctx.bf_hamming(q, t)
    oi, od = oracle_mod.bf_hamming(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    assert idx[7] == nt - 1 and dist[7] == 0 and idx[8] == 5


def test_crosscheck_golden_cv2(kctx):
    for name in ["view01_2k", "bun12_rot180_3k"]:
        g = load_golden(name)
        idx, dist, keep = kctx.bf_hamming_crosscheck(g["desc1"], g["desc2"])
        assert np.array_equal(np.nonzero(keep)[0], g["xc_query"])
        assert np.array_equal(idx[keep], g["xc_train"])


# ---- stage 2 ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["pikabun12", "disparityLR", "view01_2k", "bun12_rot180_3k"])
def test_gms_golden(ctx, name):
    g = load_golden(name)
    q = np.arange(len(g["bf_train"]), dtype=np.int32)
    for tag, rot, sc in [("00", 0, 0), ("10", 1, 0), ("01", 0, 1), ("11", 1, 1)]:
        r = ctx.gms(g["size1"], g["size2"], g["kp1"], g["kp2"], q, g["bf_train"], rot, sc)
        n = int(g["gms_len_" + tag])
        assert len(r["mask"]) == n and r["n_inliers"] == int(g["gms_n_" + tag]), tag
        assert np.array_equal(r["mask"], np.unpackbits(g["gms_mask_" + tag])[:n].astype(bool)), tag
        assert r["best_hyp"] == int(g["gms_best_" + tag]), tag
    if "anchor_default" in g:
        r = ctx.gms(g["size1"], g["size2"], g["kp1"], g["kp2"], q, g["bf_train"])
        assert r["n_inliers"] == int(g["anchor_default"])
        r = ctx.gms(g["size1"], g["size2"], g["kp1"], g["kp2"], q, g["bf_train"], True, True)
        assert r["n_inliers"] == int(g["anchor_best"]) and r["best_hyp"] == int(g["anchor_best_hyp"])


@pytest.mark.parametrize("case", sorted(MICROCASES))
def test_gms_microcases(ctx, case):
    def gms_fn(s1, s2, k1, k2, q, t, rot, sc, factor):
        return ctx.gms(s1, s2, k1, k2, q, t, rot, sc, factor)

    run_microcase(case, gms_fn)


@pytest.mark.parametrize("factor", [0.5, 3.0, 6.0, 6.0000001, 9.75])
def test_gms_threshold_factor_vs_oracle(ctx, oracle_mod, factor):
    g = load_golden("view01_2k")
    q = np.arange(len(g["bf_train"]), dtype=np.int32)
    for rot, sc in [(0, 0), (1, 1)]:
        r = ctx.gms(g["size1"], g["size2"], g["kp1"], g["kp2"], q, g["bf_train"], rot, sc, factor)
        o = oracle_mod.gms(g["size1"], g["size2"], g["kp1"], g["kp2"], q, g["bf_train"], rot, sc, factor)
        assert np.array_equal(r["mask"], o["mask"]) and r["best_hyp"] == o["best_hyp"]


def test_gms_arbitrary_matches_and_strides(ctx, oracle_mod):
    """matchGMS takes ANY match list (subset, repeated, unordered — e.g. FLANN output, DisparityUtil.cpp:143)
    and arrays of cv::KeyPoint (28 B) / cv::DMatch (16 B): exercise the strided C-ABI form."""
    import ctypes

    from sfm_gms_b200 import load_library

    g = load_golden("bun12_rot180_3k")
    rng = np.random.default_rng(4)
    n1, n2 = len(g["kp1"]), len(g["kp2"])
    sel = rng.permutation(n1)[:2000]
    qi = sel.astype(np.int32)
    ti = g["bf_train"][sel].astype(np.int32)
    qi[:50] = qi[50:100]  # repeated query indices are legal input
    o = oracle_mod.gms(g["size1"], g["size2"], g["kp1"], g["kp2"], qi, ti, True, False)
    # KeyPoint-like records: 7 floats (pt.x, pt.y, size, angle, response, octave, class_id)
    kp1 = np.zeros((n1, 7), np.float32); kp1[:, :2] = g["kp1"]; kp1[:, 2:] = 31.0
    kp2 = np.zeros((n2, 7), np.float32); kp2[:, :2] = g["kp2"]; kp2[:, 2:] = -7.0
    dm = np.zeros((len(qi), 4), np.int32); dm[:, 0] = qi; dm[:, 1] = ti; dm[:, 2] = 0; dm[:, 3] = 12345
    lib = load_library()
    mask = np.zeros(len(qi), np.uint8)
    ml, ni, bh = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib.sfmgms_gms(ctx._h, int(g["size1"][0]), int(g["size1"][1]), int(g["size2"][0]), int(g["size2"][1]),
                        kp1.ctypes.data, n1, 28, kp2.ctypes.data, n2, 28, dm.ctypes.data, dm.ctypes.data + 4, 16,
                        len(qi), 1, 0, 6.0, mask.ctypes.data, ctypes.byref(ml), ctypes.byref(ni), ctypes.byref(bh))
    assert rc == 0
    assert ml.value == len(o["mask"]) and ni.value == o["n_inliers"] and bh.value == o["best_hyp"]
    assert np.array_equal(mask[: ml.value].astype(bool), o["mask"])


def test_gms_undefined_inputs_are_errors(ctx):
    from sfm_gms_b200 import SfmGmsError

    kp = np.array([[1.0, 1.0], [639.0, 10.0]], np.float32)
    with pytest.raises(SfmGmsError) as e:
        ctx.gms((639, 480), (640, 480), kp, kp, [0, 1], [0, 1])
    assert e.value.code == 3
    with pytest.raises(SfmGmsError) as e:
        ctx.gms((640, 480), (640, 480), kp, kp, [0, 1], [0, 2])
    assert e.value.code == 4
    nan = np.array([[np.nan, 1.0], [3.0, 10.0]], np.float32)
    with pytest.raises(SfmGmsError):
        ctx.gms((640, 480), (640, 480), nan, kp, [0, 1], [0, 1])
    r = ctx.gms((640, 480), (640, 480), kp, kp, [], [])
    assert len(r["mask"]) == 0 and r["n_inliers"] == 0


# ---- fused pair, BASELINE.json configs -------------------------------------------------------------
def _oracle_pair(oracle_mod, d, rot, sc):
    idx, dist = oracle_mod.bf_hamming(d["desc1"], d["desc2"])
    q = np.arange(len(idx), dtype=np.int32)
    r = oracle_mod.gms(d["size1"], d["size2"], d["kp1"], d["kp2"], q, idx, rot, sc)
    return idx, dist, r


def test_config2_synthetic_pair(kctx, oracle_mod):
    from sfm_gms_b200 import synth

    d = synth.make_config("cfg2_640x480_10k")
    r = kctx.match_pair(d["desc1"], d["desc2"], d["kp1"], d["kp2"], d["size1"], d["size2"])
    idx, dist, o = _oracle_pair(oracle_mod, d, False, False)
    assert np.array_equal(r["train_idx"], idx) and np.array_equal(r["dist"], dist)
    assert np.array_equal(r["mask"], o["mask"]) and r["n_inliers"] == o["n_inliers"] > 1000


def test_config3_synthetic_pair_rot_scale(kctx, oracle_mod):
    from sfm_gms_b200 import synth

    d = synth.make_config("cfg3_1080p_50k_rs")
    r = kctx.match_pair(d["desc1"], d["desc2"], d["kp1"], d["kp2"], d["size1"], d["size2"], True, True)
    idx, dist, o = _oracle_pair(oracle_mod, d, True, True)
    assert np.array_equal(r["train_idx"], idx) and np.array_equal(r["dist"], dist)
    assert np.array_equal(r["mask"], o["mask"]) and r["best_hyp"] == o["best_hyp"] and o["best_hyp"] not in (0, -1)


def test_config4_synthetic_200k(kctx, oracle_mod):
    """Full size 200k x 200k.  The oracle BF would take minutes, so: (a) a 1,500-row sample of the queries is
    checked bit-exactly against the oracle over the FULL train set, (b) the distance of every returned match
    is recomputed independently (numpy) and must be consistent, (c) GMS on the GPU's full match list is
    compared bit-exactly with the oracle GMS on the same list (H=1 and H=40)."""
    from sfm_gms_b200 import synth

    d = synth.make_config("cfg4_4k_200k")
    n = len(d["desc1"])
    r = kctx.match_pair(d["desc1"], d["desc2"], d["kp1"], d["kp2"], d["size1"], d["size2"])
    rng = np.random.default_rng(44)
    sel = np.sort(rng.permutation(n)[:1500])
    oi, od = oracle_mod.bf_hamming(d["desc1"][sel], d["desc2"])
    assert np.array_equal(r["train_idx"][sel], oi) and np.array_equal(r["dist"][sel], od)
    x = d["desc1"] ^ d["desc2"][r["train_idx"]]
    assert np.array_equal(np.unpackbits(x, axis=1).sum(1), r["dist"])
    q = np.arange(n, dtype=np.int32)
    o = oracle_mod.gms(d["size1"], d["size2"], d["kp1"], d["kp2"], q, r["train_idx"])
    assert np.array_equal(r["mask"], o["mask"]) and r["n_inliers"] == o["n_inliers"]
    r40 = kctx.gms(d["size1"], d["size2"], d["kp1"], d["kp2"], q, r["train_idx"], True, True)
    o40 = oracle_mod.gms(d["size1"], d["size2"], d["kp1"], d["kp2"], q, r["train_idx"], True, True)
    assert np.array_equal(r40["mask"], o40["mask"]) and r40["best_hyp"] == o40["best_hyp"]


# ---- multi-pair ------------------------------------------------------------------------------------
def _ragged_set(rng, sizes_n):
    from sfm_gms_b200 import synth

    descs, kps, wh = [], [], []
    base = synth.make_pair(800, 600, max(sizes_n), 123)
    for k, n in enumerate(sizes_n):
        w, h = 800 - 16 * k, 600 + 8 * k
        idx = rng.permutation(len(base["desc1"]))[:n]
        descs.append(synth.flip_bits(rng, base["desc1"][idx], 0.04))
        xy = base["kp1"][idx] * np.array([w / 800.0, h / 600.0], np.float32) + rng.normal(0, 0.7, (n, 2))
        kps.append(synth._inside(xy, w, h))
        wh.append((w, h))
    off = np.concatenate([[0], np.cumsum(sizes_n)]).astype(np.int64)
    return off, np.concatenate(descs), np.concatenate(kps), np.array(wh, np.int32)


@pytest.mark.parametrize("chunk_mb", [8, 64])
@pytest.mark.parametrize("rot,sc", [(0, 0), (1, 1)])
def test_match_pairs_ragged_vs_oracle(kctx, oracle_mod, rot, sc, chunk_mb):
    from sfm_gms_b200 import api

    rng = np.random.default_rng(77)
    sizes_n = [3000, 2500, 0, 1777, 4096, 1]
    off, desc, kp, wh = _ragged_set(rng, sizes_n)
    kctx.set_images(off, desc, kp, wh)
    pairs = np.array([(0, 1), (1, 0), (0, 3), (4, 0), (2, 1), (1, 2), (5, 4), (3, 3), (4, 5)], np.int32)
    kctx.set_option(api.OPT_GMS_CHUNK_BYTES, chunk_mb << 20)  # 8 MB: several GMS chunks; 64 MB: one chunk that
    # contains pairs with an empty train image (rows without matches) in its middle
    try:
        out = kctx.match_pairs(pairs, rot, sc)
    finally:
        kctx.set_option(api.OPT_GMS_CHUNK_BYTES, 64 << 20)
    mo = out["offsets"]
    for p, (a, b) in enumerate(pairs):
        d1, d2 = desc[off[a]:off[a + 1]], desc[off[b]:off[b + 1]]
        k1, k2 = kp[off[a]:off[a + 1]], kp[off[b]:off[b + 1]]
        oi, od = oracle_mod.bf_hamming(d1, d2)
        sl = slice(mo[p], mo[p + 1])
        if len(d2) == 0:
            assert (out["train_idx"][sl] == -1).all() and out["n_inliers"][p] == 0 and out["mask_len"][p] == 0
            continue
        assert np.array_equal(out["train_idx"][sl], oi) and np.array_equal(out["dist"][sl], od), p
        o = oracle_mod.gms(wh[a], wh[b], k1, k2, np.arange(len(oi), dtype=np.int32), oi, rot, sc)
        assert out["n_inliers"][p] == o["n_inliers"] and out["mask_len"][p] == len(o["mask"]), p
        assert out["best_hyp"][p] == o["best_hyp"], p
        m = out["mask"][sl].astype(bool)
        assert np.array_equal(m[: len(o["mask"])], o["mask"]) and not m[len(o["mask"]):].any(), p
        # §8f-1: inlier coordinate compaction (SfMUtil.cpp:25-35)
        p1, p2, n = kctx.inlier_points(p, len(oi) + 1)
        keep = np.nonzero(o["mask"])[0] if len(o["mask"]) else np.zeros(0, int)
        assert n == len(keep)
        assert np.array_equal(p1, k1[keep]) and np.array_equal(p2, k2[oi[keep]])


@pytest.mark.parametrize("rot,sc", [(0, 0), (1, 0)])
def test_match_image_set_pipelined_equals_match_pairs(kctx, rot, sc):
    """The internally pipelined one-call form (chunked H2D / compute / D2H overlap) must return exactly what
    set_images + match_pairs returns, for in-order, out-of-order and repeated pairs over a ragged set."""
    rng = np.random.default_rng(78)
    sizes_n = [2100, 1900, 0, 1777, 2500, 1, 3000, 2048, 2047]
    off, desc, kp, wh = _ragged_set(rng, sizes_n)
    pairs = np.array([(0, 1), (1, 3), (3, 4), (4, 6), (6, 7), (7, 8), (8, 0), (2, 1), (1, 2), (5, 4), (0, 1), (7, 6),
                      (6, 4), (3, 3), (4, 8), (8, 7), (0, 8), (1, 0)], np.int32)
    kctx.set_images(off, desc, kp, wh)
    a = kctx.match_pairs(pairs, rot, sc)
    b = kctx.match_image_set(off, desc, kp, wh, pairs, rot, sc)
    for k in ["n_inliers", "best_hyp", "mask_len", "offsets", "train_idx", "dist", "mask"]:
        assert np.array_equal(a[k], b[k]), k
    c = kctx.match_pairs(pairs, rot, sc)   # the set stays registered after the pipelined call
    assert np.array_equal(a["mask"], c["mask"]) and np.array_equal(a["train_idx"], c["train_idx"])


def test_match_pairs_properties_full_batch(kctx):
    """Size-independent properties on a config-2-sized batch: determinism (two runs identical), symmetry of
    the distance (dist(i -> j*) equals the recomputed popcount), and n_inliers == popcount(mask)."""
    from sfm_gms_b200 import synth

    s = synth.make_pair_batch(6)
    kctx.set_images(s["offsets"], s["desc"], s["kp"], s["sizes"])
    a = kctx.match_pairs(s["pairs"])
    b = kctx.match_pairs(s["pairs"])
    for k in ["train_idx", "dist", "mask", "n_inliers"]:
        assert np.array_equal(a[k], b[k]), k
    mo = a["offsets"]
    for p, (i, j) in enumerate(s["pairs"]):
        sl = slice(mo[p], mo[p + 1])
        d1 = s["desc"][s["offsets"][i]:s["offsets"][i + 1]]
        d2 = s["desc"][s["offsets"][j]:s["offsets"][j + 1]]
        x = d1 ^ d2[a["train_idx"][sl]]
        assert np.array_equal(np.unpackbits(x, axis=1).sum(1), a["dist"][sl])
        assert a["n_inliers"][p] == int(a["mask"][sl].sum()) > 1000


# ---- cv2-shaped Python surface -----------------------------------------------------------------------
def test_cv2_shaped_api(ctx):
    import sfm_gms_b200 as sg

    g = load_golden("view01_2k")
    m = sg.BFMatcher(sg.NORM_HAMMING).match(g["desc1"], g["desc2"])
    assert [x.trainIdx for x in m] == g["bf_train"].tolist() and m[5].queryIdx == 5 and m[5].imgIdx == 0
    assert [int(x.distance) for x in m] == g["bf_dist"].tolist()
    out = sg.matchGMS(g["size1"], g["size2"], g["kp1"], g["kp2"], m, withRotation=False, withScale=False)
    exp = np.unpackbits(g["gms_mask_00"])[: len(m)].astype(bool)
    assert [x.queryIdx for x in out] == np.nonzero(exp)[0].tolist()
    # upstream class: scale FIRST, rotation second
    n, mask = sg.gms_matcher(g["kp1"], g["size1"], g["kp2"], g["size2"], m).GetInlierMask(False, True)
    assert n == int(g["gms_n_10"]) and np.array_equal(mask, np.unpackbits(g["gms_mask_10"])[: len(m)].astype(bool))
    xm = sg.BFMatcher(sg.NORM_HAMMING, crossCheck=True).match(g["desc1"], g["desc2"])
    assert [x.queryIdx for x in xm] == g["xc_query"].tolist() and [x.trainIdx for x in xm] == g["xc_train"].tolist()


# ---- the C++ shim (sfmgms.hpp): BFMatcher::match + matchGMS exactly as FeatureMatchUtil.cpp:66-69 calls them ----
@pytest.mark.parametrize("rot,sc,tag", [(0, 0, "00"), (1, 1, "11")])
def test_cxx_shim_demo_matches_golden(tmp_path, oracle_mod, rot, sc, tag):
    import os
    import subprocess

    from conftest import ROOT

    exe = os.path.join(ROOT, "sfm_gms_b200", "cxx", "demo_match")
    assert os.path.exists(exe), "build it with __graft_entry__.build()"
    g = load_golden("bun12_rot180_3k")
    n1, n2 = len(g["kp1"]), len(g["kp2"])
    pair = tmp_path / "pair.bin"
    with open(pair, "wb") as f:
        np.array([g["size1"][0], g["size1"][1], g["size2"][0], g["size2"][1], n1, n2], np.int32).tofile(f)
        g["kp1"].astype(np.float32).tofile(f)
        g["kp2"].astype(np.float32).tofile(f)
        g["desc1"].tofile(f)
        g["desc2"].tofile(f)
    out = tmp_path / "out.bin"
    subprocess.check_call([exe, str(pair), str(rot), str(sc), str(out)], timeout=120)
    raw = np.fromfile(out, np.uint8)
    n = int(raw[:4].view(np.int32)[0])
    o = 4
    ti = raw[o:o + 4 * n].view(np.int32); o += 4 * n
    di = raw[o:o + 4 * n].view(np.int32); o += 4 * n
    ng = int(raw[o:o + 4].view(np.int32)[0]); o += 4
    gq = raw[o:o + 4 * ng].view(np.int32); o += 4 * ng
    nc = int(raw[o:o + 4].view(np.int32)[0]); o += 4
    mk = raw[o:o + n].astype(bool); o += n
    nf = int(raw[o:o + 4].view(np.int32)[0]); o += 4
    nb = int(raw[o:o + 4].view(np.int32)[0]); o += 4
    bq = raw[o:o + 4 * nb].view(np.int32); o += 4 * nb
    bt = raw[o:o + 4 * nb].view(np.int32)
    eq, et, _ = oracle_mod.brute_force_match(g["desc1"], g["desc2"], "hamming", True, 4.0, 500)
    assert np.array_equal(bq, eq) and np.array_equal(bt, et)          # sfmgms::bruteForceMatch (C++ shim)
    assert n == n1 and np.array_equal(ti, g["bf_train"]) and np.array_equal(di, g["bf_dist"])
    exp = np.unpackbits(g["gms_mask_" + tag])[:n].astype(bool)
    assert ng == int(g["gms_n_" + tag]) and np.array_equal(gq, np.nonzero(exp)[0])
    assert nc == ng and np.array_equal(mk, exp) and nf == ng


# ---- (§8f-3) the reference's literal main path: SIFT + BFMatcher(NORM_L2) + matchGMS(true, true) ----------------
@pytest.fixture(params=["dp4a", "tcgen05"])
def l2ctx(ctx, request):
    from sfm_gms_b200 import api

    ctx.set_option(api.OPT_L2_KERNEL, 1 if request.param == "dp4a" else 2)
    yield ctx
    ctx.set_option(api.OPT_L2_KERNEL, 0)


def test_sift_l2_pipeline_golden_cv2(l2ctx, oracle_mod):
    import sfm_gms_b200 as sg

    ctx = l2ctx

    g = load_golden("sift_view01_1500")
    d1, d2 = g["desc1"].astype(np.float32), g["desc2"].astype(np.float32)
    idx, dist = ctx.bf_l2(d1, d2)
    assert np.array_equal(idx, g["l2_train"]) and np.array_equal(dist, g["l2_dist"])   # bit-exact float distances
    m = sg.BFMatcher(sg.NORM_L2).match(d1, d2)
    out = sg.matchGMS(g["size1"], g["size2"], g["kp1"], g["kp2"], m, withRotation=True, withScale=True)
    n = int(g["gms_len_11"])
    exp = np.unpackbits(g["gms_mask_11"])[:n].astype(bool)
    assert [x.queryIdx for x in out] == np.nonzero(exp)[0].tolist() and len(out) == int(g["gms_n_11"])
    # float-tie: sqrtf(n) == sqrtf(n+1) above 2^22 -> the LOWER index wins although its exact d2 is larger
    idx, dist = ctx.bf_l2(g["tie_q"], g["tie_t"])
    assert idx[0] == 0 and dist[0] == g["tie_dist"]


def test_l2_ragged_vs_oracle_and_errors(l2ctx, oracle_mod):
    from sfm_gms_b200 import SfmGmsError

    ctx = l2ctx

    rng = np.random.default_rng(21)
    for nq, nt in [(1, 1), (3, 700), (513, 257), (2000, 1999), (385, 241), (4000, 5000)]:
        hi = 256 if nq % 2 else 40            # full-range values: squared distances above 2^22 (float-tie fixup path)
        q = rng.integers(0, hi, (nq, 128)).astype(np.float32)
        t = rng.integers(0, hi, (nt, 128)).astype(np.float32)
        if nt > 4:
            t[nt // 2] = t[1]; q[0] = t[1]
        idx, dist = ctx.bf_l2(q, t)
        oi, od = oracle_mod.bf_l2(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    idx, dist = ctx.bf_l2(q, np.zeros((0, 128), np.float32))
    assert len(idx) == 0
    with pytest.raises(SfmGmsError):      # OpenCV asserts equal widths
        ctx.bf_l2(q[:, :64], t)
    with pytest.raises(SfmGmsError):
        ctx.bf_l2(np.zeros((4, 300), np.float32), np.zeros((4, 300), np.float32))   # dim > 256


def _rootsift(g):
    d1, d2 = g["desc1"].astype(np.float32), g["desc2"].astype(np.float32)
    r1 = np.sqrt(d1 / np.maximum(d1.sum(1, keepdims=True), 1e-7)).astype(np.float32)
    r2 = np.sqrt(d2 / np.maximum(d2.sum(1, keepdims=True), 1e-7)).astype(np.float32)
    return {"root128": (r1, r2), "root64": (np.ascontiguousarray(r1[:, :64]), np.ascontiguousarray(r2[:, :64])),
            "root70": (np.ascontiguousarray(r1[:600, 3:73]), np.ascontiguousarray(r2[:900, 3:73]))}


def test_l2_general_float_golden_cv2(l2ctx):
    """(§8f-3) non-integer float descriptors (RootSIFT; widths 128, 64 and 70 = scalar tail): trainIdx AND the float
    distances bit-identical to cv2 -- the fp32 kernel restates OpenCV's normL2Sqr_ addition order."""
    import sfm_gms_b200 as sg

    ctx = l2ctx
    g = load_golden("l2_float")
    data = _rootsift(load_golden("sift_view01_1500"))
    for name, (a, b) in data.items():
        idx, dist = ctx.bf_l2(a, b)
        assert np.array_equal(idx, g[name + "_train"]) and np.array_equal(dist, g[name + "_dist"]), name
    a, b = data["root128"]
    m = sg.BFMatcher(sg.NORM_L2, True, ctx=ctx).match(a, b)
    assert np.array_equal(np.array([(x.queryIdx, x.trainIdx) for x in m], np.int32), g["root128_xc"])
    assert np.array_equal(np.array([x.distance for x in m], np.float32), g["root128_xc_dist"])


def test_l2_general_float_vs_oracle_ragged(ctx, oracle_mod):
    from sfm_gms_b200 import api

    rng = np.random.default_rng(77)
    for nq, nt, dim in [(1, 1, 128), (3, 700, 128), (513, 257, 64), (130, 1999, 33), (385, 241, 7), (64, 64, 16),
                        (1000, 1500, 256), (200, 300, 1)]:
        q = (rng.standard_normal((nq, dim)) * 11).astype(np.float32)
        t = (rng.standard_normal((nt, dim)) * 11).astype(np.float32)
        if nt > 4:
            t[nt // 2] = t[1]; q[0] = t[1]          # exact ties: lowest train index must win
        idx, dist = ctx.bf_l2(q, t)
        oi, od = oracle_mod.bf_l2(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt, dim)
    # forcing the fp32 kernel on integer-valued SIFT data gives the same bits as the tensor-core path
    g = load_golden("sift_view01_1500")
    ctx.set_option(api.OPT_L2_KERNEL, 3)
    try:
        idx, dist = ctx.bf_l2(g["desc1"].astype(np.float32), g["desc2"].astype(np.float32))
    finally:
        ctx.set_option(api.OPT_L2_KERNEL, 0)
    assert np.array_equal(idx, g["l2_train"]) and np.array_equal(dist, g["l2_dist"])


# ---- (§8f-2) bruteForceMatch: BFMatcher(NORM_L2, crossCheck=true) + sort + ratio prune + cap (FeatureMatchUtil.cpp:20-31)
BF_CASES = [("l2", "sift_view01_1500", False), ("l2_rev", "sift_view01_1500", True), ("ham", "view01_2k", False)]


def _bruteforce_inputs(name, src, rev):
    g = load_golden(src)
    if name == "ham":
        return g["desc1"], g["desc2"]
    d1, d2 = g["desc1"].astype(np.float32), g["desc2"].astype(np.float32)
    return (d2[:700], d1) if rev else (d1, d2)


@pytest.mark.parametrize("name,src,rev", BF_CASES)
def test_bruteforce_golden_cv2(l2ctx, name, src, rev):
    import sfm_gms_b200 as sg

    ctx = l2ctx
    g = load_golden("bruteforce")
    q, t = _bruteforce_inputs(name, src, rev)
    norm = sg.NORM_HAMMING if name == "ham" else sg.NORM_L2
    # the matcher bruteForceMatch constructs: cv2-pinned (queryIdx, trainIdx, distance) list
    m = sg.BFMatcher(norm, True, ctx=ctx).match(q, t)
    full = g[name + "_xc_full"]
    assert [(x.queryIdx, x.trainIdx) for x in m] == [tuple(r) for r in full.tolist()]
    assert np.array_equal(np.array([x.distance for x in m], np.float32), g[name + "_xc_full_dist"])
    for xc in (True, False):
        for coef, cap in ((4.0, 500), (1.5, 100000), (4.0, 37)):
            key = "%s_%s_c%g_m%d" % (name, "xc" if xc else "nn", coef, cap)
            qi, ti, d = ctx.brute_force_match(q, t, norm, xc, coef, cap)
            assert np.array_equal(np.stack([qi, ti], 1).reshape(-1, 2), g[key]) and np.array_equal(d, g[key + "_dist"]), key
    # the reference's own entry points, default constants (kDistanceCoef = 4, kMaxMatchingSize = 500)
    out = sg.bruteForceMatch(q, t, ctx=ctx)
    key = name + "_xc_c4_m500"
    assert [(x.queryIdx, x.trainIdx) for x in out] == [tuple(r) for r in g[key].tolist()]
    from sfm_gms_b200 import api
    out = api.match(q, t, 1.5, 100000, ctx=ctx)
    assert [(x.queryIdx, x.trainIdx) for x in out] == [tuple(r) for r in g[name + "_nn_c1.5_m100000"].tolist()]


def test_l2_crosscheck_vs_oracle_ragged(l2ctx, oracle_mod):
    ctx = l2ctx
    rng = np.random.default_rng(33)
    for nq, nt in [(1, 1), (5, 300), (700, 129), (1500, 1501)]:
        q = rng.integers(0, 60, (nq, 128)).astype(np.float32)
        t = rng.integers(0, 60, (nt, 128)).astype(np.float32)
        k = min(nq, nt) // 2
        t[:k] = q[:k]                       # mutual pairs ...
        if nt > 3 and nq > 3:
            t[nt - 1] = q[0]; q[nq - 1] = q[1]   # ... and duplicates that the tie rules must break the same way
        idx, dist, keep = ctx.bf_l2_crosscheck(q, t)
        oi, od, ok = oracle_mod.bf_l2_crosscheck(q, t)
        assert np.array_equal(keep, ok) and np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt)
        a = ctx.brute_force_match(q, t, 4, True, 4.0, 500)
        b = oracle_mod.brute_force_match(q, t, "l2", True, 4.0, 500)
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), (nq, nt)
    qi, ti, d = ctx.brute_force_match(q, np.zeros((0, 128), np.float32))
    assert len(qi) == 0


# ---- (§8f-4, first half) ORB descriptors on provided level-0 keypoints (DisparityUtil.cpp:107, 127-134) ----------
@pytest.mark.parametrize("name", ["view0_bgr", "disp_gray"])
def test_orb_compute_golden_cv2(ctx, name):
    import sfm_gms_b200 as sg

    g = load_golden("orb_compute")
    kept, desc = ctx.orb_compute(g[name + "_img"], g[name + "_pts"], g[name + "_ang"])
    assert np.array_equal(kept, g[name + "_kept"]) and np.array_equal(desc, g[name + "_desc"])   # every bit, vs cv2

    class KP:                                   # cv2.KeyPoint-shaped objects through the ORB look-alike
        def __init__(self, p, a):
            self.pt, self.angle, self.octave = (float(p[0]), float(p[1])), float(a), 0

    kps = [KP(p, a) for p, a in zip(g[name + "_pts"][:400], g[name + "_ang"][:400])]
    out_k, out_d = sg.ORB_create(ctx=ctx).compute(g[name + "_img"], kps)
    n = int((g[name + "_kept"] < 400).sum())
    assert len(out_k) == n and np.array_equal(out_d, g[name + "_desc"][:n])
    assert [kps.index(k) for k in out_k] == g[name + "_kept"][:n].tolist()


def test_orb_compute_vs_oracle_edges(ctx):
    from oracle import orb
    from sfm_gms_b200 import SfmGmsError

    rng = np.random.default_rng(5)
    for (h, w, ch) in [(63, 63, 1), (64, 70, 3), (97, 65, 1), (200, 333, 3)]:
        img = rng.integers(0, 256, (h, w) if ch == 1 else (h, w, 3), dtype=np.uint8)
        n = 500
        pts = np.stack([rng.uniform(-2, w + 2, n), rng.uniform(-2, h + 2, n)], 1).astype(np.float32)
        pts[:40] = np.round(pts[:40]) + 0.5                       # half-way positions: cvRound is half-to-even
        ang = rng.uniform(-10, 370, n).astype(np.float32)
        kept, desc = ctx.orb_compute(img, pts, ang)
        ok, od = orb.orb_compute(img, pts, ang)
        assert np.array_equal(kept, ok) and np.array_equal(desc, od), (h, w, ch)
    kept, desc = ctx.orb_compute(img, np.zeros((0, 2), np.float32))
    assert len(kept) == 0 and desc.shape == (0, 32)
    with pytest.raises(SfmGmsError):
        ctx.orb_compute(img, pts, ang, octaves=np.full(n, 16, np.int32))      # octaves above 15 are refused


def test_orb_compute_octaves_golden_cv2(ctx):
    """provided keypoints on pyramid levels 0..3, not sorted by octave: OpenCV regroups them level by level and
    samples coarse-level patches that may stick out of the level into its reflected border"""
    from oracle import orb

    g = load_golden("orb_compute_octaves")
    kept, desc = ctx.orb_compute(g["img"], g["pts"], g["ang"], g["oct"])
    assert np.array_equal(kept, g["kept"]) and np.array_equal(desc, g["desc"])
    o = np.sort(g["oct"])                                          # sorted input: order kept as given
    kept, desc = ctx.orb_compute(g["img"], g["pts"], g["ang"], o)
    ok, od = orb.orb_compute(g["img"], g["pts"], g["ang"], o)
    assert np.array_equal(kept, ok) and np.array_equal(desc, od)


@pytest.mark.parametrize("name", ["view0_bgr", "pika_gray"])
@pytest.mark.parametrize("nf,thr", [(500, 20), (3000, 0)])
def test_orb_detect_and_compute_golden_cv2(ctx, name, nf, thr):
    """(§8f-4) ORB::create(nf) + detectAndCompute: the keypoint LIST (order, pt, size, angle, response, octave) and
    every descriptor bit equal cv2 (DisparityUtil.cpp:107,139-140 defaults; BASELINE config 1's generator settings)."""
    import sfm_gms_b200 as sg

    g = load_golden("orb_detect")
    key = "%s_%d_%d" % (name, nf, thr)
    kp, desc = ctx.orb_detect_and_compute(g[name + "_img"], nf, thr)
    assert kp.shape == g[key + "_kp"].shape and np.array_equal(kp, g[key + "_kp"])
    assert np.array_equal(desc, g[key + "_desc"])
    orb = sg.ORB_create(nf, ctx=ctx)
    orb.setFastThreshold(thr)
    kps = orb.detect(g[name + "_img"])
    assert len(kps) == len(kp) and kps[0].pt == (float(kp[0, 0]), float(kp[0, 1])) and kps[-1].octave == int(kp[-1, 5])


def test_orb_detect_vs_oracle_random_images(ctx):
    from oracle import orb

    rng = np.random.default_rng(9)
    for (h, w, ch, nf, thr) in [(240, 320, 1, 300, 20), (200, 260, 3, 800, 5), (150, 150, 1, 100, 40), (90, 400, 1, 50, 10)]:
        base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2) if ch == 1 else (h // 8 + 2, w // 8 + 2, 3), dtype=np.uint8)
        img = np.kron(base, np.ones((8, 8) if ch == 1 else (8, 8, 1), np.uint8))[:h, :w].copy()   # blocky: many corners, many ties
        img = (img.astype(np.int32) + rng.integers(-6, 7, img.shape)).clip(0, 255).astype(np.uint8)
        kp, desc = ctx.orb_detect_and_compute(img, nf, thr)
        okp, od = orb.orb_detect_and_compute(img, nf, thr)
        assert kp.shape == okp.shape and np.array_equal(kp.astype(np.float64), okp), (h, w, ch, nf, thr)
        assert np.array_equal(desc, od), (h, w, ch, nf, thr)
    kp, desc = ctx.orb_detect_and_compute(np.full((100, 100), 7, np.uint8))      # nothing to detect
    assert kp.shape == (0, 6) and desc.shape == (0, 32)


def test_full_pipeline_image_to_inliers(ctx, oracle_mod):
    """image pair -> ORB -> BF-Hamming -> GMS, the whole chain on the device, against the whole chain of oracles:
    DisparityUtil.cpp:107,139-149 with the brute-force matcher of FeatureMatchUtil.cpp:66-69."""
    from oracle import orb as orb_oracle

    g = load_golden("orb_detect")
    img1 = g["view0_bgr_img"]
    img2 = np.ascontiguousarray(img1[7:, 11:])               # a translated view of the same scene
    out = []
    for fn, bf, gms in ((ctx.orb_detect_and_compute, ctx.bf_hamming, ctx.gms),
                        (orb_oracle.orb_detect_and_compute, oracle_mod.bf_hamming, oracle_mod.gms)):
        k1, d1 = fn(img1, 1500, 5)
        k2, d2 = fn(img2, 1500, 5)
        idx, dist = bf(d1, d2)
        s1, s2 = (img1.shape[1], img1.shape[0]), (img2.shape[1], img2.shape[0])
        r = gms(s1, s2, np.asarray(k1)[:, :2].astype(np.float32), np.asarray(k2)[:, :2].astype(np.float32),
                np.arange(len(idx), dtype=np.int32), idx, False, False)
        out.append((np.asarray(k1, np.float64), d1, idx, dist, np.asarray(r["mask"]), r["n_inliers"]))
    a, b = out
    assert all(np.array_equal(x, y) for x, y in zip(a[:5], b[:5])) and a[5] == b[5]
    assert a[5] > 200                                         # the translation is found: most matches are inliers


def test_orb_detect_and_compute_vs_live_cv2_full_size(ctx):
    """when cv2 is importable: a full-size (1390 x 1110) tie-heavy image, 10,000 features, FAST threshold 0 -- the
    keypoint list and all 2.56 M descriptor bits equal OpenCV's"""
    cv2 = pytest.importorskip("cv2")
    g = load_golden("orb_detect")
    img = np.tile(g["view0_bgr_img"], (4, 4, 1))[:1110, :1390].copy()
    kp, desc = ctx.orb_detect_and_compute(img, 10000, 0)
    orb = cv2.ORB_create(10000)
    orb.setFastThreshold(0)
    ref_k, ref_d = orb.detectAndCompute(img, None)
    ref = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in ref_k], np.float32)
    assert kp.shape == ref.shape and np.array_equal(kp, ref) and np.array_equal(desc, ref_d)


def test_new_entry_points_report_errors(ctx):
    """error behaviour of the §8f entry points through the raw C ABI (no exceptions across the boundary)"""
    import ctypes

    lib, h = ctx._lib, ctx._h
    g = load_golden("orb_detect")
    img = np.ascontiguousarray(g["pika_gray_img"])
    hh, ww = img.shape
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rec = np.zeros((8, 7), np.float32)
    desc = np.zeros((8, 32), np.uint8)
    n = ctypes.c_int(0)
    # capacity too small: SFMGMS_ERR_ARG and the needed count comes back
    rc = lib.sfmgms_orb_detect_and_compute(h, p(img), ww, hh, 1, ww, 500, 20, p(rec), p(desc), 8, ctypes.byref(n))
    assert rc == 1 and n.value == len(g["pika_gray_500_20_kp"]) and b"capacity" in lib.sfmgms_last_error(h)
    assert lib.sfmgms_orb_detect_and_compute(h, p(img), ww, hh, 2, 2 * ww, 500, 20, p(rec), p(desc), 8, ctypes.byref(n)) == 1
    assert lib.sfmgms_orb_detect_and_compute(h, p(img), ww, hh, 1, ww - 1, 500, 20, p(rec), p(desc), 8, ctypes.byref(n)) == 1
    assert lib.sfmgms_orb_detect_and_compute(h, None, ww, hh, 1, ww, 500, 20, p(rec), p(desc), 8, ctypes.byref(n)) == 1
    # detect only (descriptors = NULL) is allowed
    big = np.zeros((2000, 7), np.float32)
    assert lib.sfmgms_orb_detect_and_compute(h, p(img), ww, hh, 1, ww, 500, 20, p(big), None, 2000, ctypes.byref(n)) == 0
    assert n.value == len(g["pika_gray_500_20_kp"]) and np.array_equal(big[: n.value, :5], g["pika_gray_500_20_kp"][:, :5])
    # bruteForceMatch: unknown norm, capacity too small
    q = np.zeros((4, 32), np.uint8)
    out = np.zeros(4, np.int32)
    d = np.zeros(4, np.float32)
    assert lib.sfmgms_brute_force_match(h, 5, 1, p(q), 4, p(q), 4, 32, 4.0, 500, p(out), p(out), p(d), 4, ctypes.byref(n)) == 1
    assert lib.sfmgms_brute_force_match(h, 6, 0, p(q), 4, p(q), 4, 32, 4.0, 500, p(out), p(out), p(d), 2, ctypes.byref(n)) == 1
    assert lib.sfmgms_brute_force_match(h, 6, 0, p(q), 4, p(q), 4, 32, 4.0, 500, p(out), p(out), p(d), 4, ctypes.byref(n)) == 0 and n.value == 4


ORB_PARAM_SETS = [(800, 1.5, 4, 31, 0, 10), (300, 1.1, 12, 19, 0, 20), (700, 1.2, 8, 31, 1, 20), (400, 1.3, 1, 25, 1, 5),
                  (2000, 2.0, 3, 40, 0, 0), (500, 1.2, 8, 31, 0, 20, 3, 31), (500, 1.2, 8, 31, 0, 20, 4, 31), (600, 1.2, 6, 31, 0, 10, 2, 21),
                  (600, 1.3, 5, 25, 1, 10, 3, 41), (800, 1.2, 8, 5, 0, 20, 2, 31), (800, 1.2, 4, 0, 0, 5, 4, 15), (300, 1.2, 8, 12, 0, 20, 2, 31)]


@pytest.mark.parametrize("i", range(len(ORB_PARAM_SETS)))
def test_orb_non_default_parameters_golden_cv2(ctx, i):
    """ORB::create(nfeatures, scaleFactor, nlevels, edgeThreshold, 0, WTA_K, scoreType, patchSize, fastThreshold):
    pyramids of 1-12 levels, scale factors 1.1-2.0, FAST_SCORE ranking, border widths 0-40 (reads that leave the
    level), WTA_K 3 and 4, random patterns for other patch sizes -- keypoint list and descriptors equal cv2"""
    import sfm_gms_b200 as sg
    from sfm_gms_b200 import SfmGmsError

    g = load_golden("orb_detect")
    nf, sf, nl, edge, score, thr = ORB_PARAM_SETS[i][:6]
    wta, patch = ORB_PARAM_SETS[i][6:] if len(ORB_PARAM_SETS[i]) > 6 else (2, 31)
    orb = sg.ORB_create(nf, sf, nl, edge, 0, wta, score, patch, thr, ctx=ctx)
    kps, desc = orb.detectAndCompute(g["view0_bgr_img"])
    ref = g["params%d_kp" % i]
    got = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in kps], np.float32).reshape(-1, 6)
    assert np.array_equal(got, ref) and np.array_equal(desc, g["params%d_desc" % i])
    if i == 0:
        for bad in (dict(WTA_K=5), dict(firstLevel=1), dict(patchSize=64), dict(edgeThreshold=-1), dict(scaleFactor=1.0),
                    dict(nlevels=0), dict(scoreType=2)):
            with pytest.raises(SfmGmsError):                   # refused loudly, never approximated
                sg.ORB_create(100, ctx=ctx, **bad).detectAndCompute(g["view0_bgr_img"])


def test_device_descriptors_must_be_16_byte_aligned(ctx):
    """caller-owned DEVICE arrays: the documented alignment (16 bytes for descriptors) is enforced, not assumed"""
    torch = pytest.importorskip("torch")
    from sfm_gms_b200 import SfmGmsError, api

    dev = torch.device("cuda", 0)
    raw = torch.zeros(100 * 32 + 4, dtype=torch.uint8, device=dev)
    kp = torch.zeros(100, 2, dtype=torch.float32, device=dev)
    off = np.array([0, 60, 100], np.int64)
    sizes = np.tile(np.array([[320, 240]], np.int32), (2, 1))
    with pytest.raises(SfmGmsError):
        ctx.set_images_raw(off, raw.data_ptr() + 4, kp.data_ptr(), sizes, api.SFMGMS_DEVICE, keepalive=(raw, kp))
    ctx.set_images_raw(off, raw.data_ptr(), kp.data_ptr(), sizes, api.SFMGMS_DEVICE, keepalive=(raw, kp))     # aligned: accepted


def test_pixels_to_inlier_matches_api(ctx, oracle_mod):
    """sfmgms_set_images_from_pixels + sfmgms_match_pairs: three views of a scene in, inlier matches of all ordered
    pairs out -- against ORB oracle -> BF oracle -> GMS oracle on the same pixels"""
    from oracle import orb as orb_oracle

    g = load_golden("orb_detect")
    base = g["view0_bgr_img"]
    imgs = [base, np.ascontiguousarray(base[7:, 11:]), np.ascontiguousarray(base[:-9, 5:-3])]
    off = ctx.set_images_from_pixels(imgs, nfeatures=1200, fast_threshold=5)
    exp = [orb_oracle.orb_detect_and_compute(im, 1200, 5) for im in imgs]
    assert off.tolist() == np.concatenate([[0], np.cumsum([len(k) for k, _ in exp])]).tolist()
    for i, (k, _) in enumerate(exp):
        assert np.array_equal(ctx.get_image_keypoints(i).astype(np.float64), k)
    pairs = np.array([[0, 1], [1, 2], [2, 0], [1, 0]], np.int32)
    r = ctx.match_pairs(pairs, False, False)
    for p, (a, b) in enumerate(pairs):
        (ka, da), (kb, db) = exp[a], exp[b]
        idx, dist = oracle_mod.bf_hamming(da, db)
        sa, sb = (imgs[a].shape[1], imgs[a].shape[0]), (imgs[b].shape[1], imgs[b].shape[0])
        o = oracle_mod.gms(sa, sb, ka[:, :2].astype(np.float32), kb[:, :2].astype(np.float32),
                           np.arange(len(idx), dtype=np.int32), idx, False, False)
        lo, hi = r["offsets"][p], r["offsets"][p + 1]
        assert np.array_equal(r["train_idx"][lo:hi], idx) and np.array_equal(r["dist"][lo:hi], dist)
        assert np.array_equal(r["mask"][lo:hi].astype(bool), o["mask"]) and r["n_inliers"][p] == o["n_inliers"] > 100
