"""Stage 2 (GMS) against the reference's OWN MACHINE CODE.

tests/golden/gms_dll.npz holds what /root/reference/SfM-GMS/bin/opencv_xfeatures2d452.dll itself computed
(exported cv::xfeatures2d::matchGMS, GMSMatcher::run per hypothesis, getGridIndexLeft/Right, ROT/SCALE tables),
recorded by tests/golden/make_gms_dll_golden.py through the PE host in oracle/dllref/.  Here:
  * CPU  (-m "not gpu"): the oracle restatement == the DLL, on every recorded case;
  * CPU, build container only: live randomised comparison oracle == DLL (skipped where the DLL is absent);
  * GPU  (-m gpu): the CUDA path through the C ABI == the DLL.
Inputs are rebuilt by tests/gms_dll_cases.py.
"""
import numpy as np
import pytest

import gms_dll_cases as C
from conftest import load_golden


@pytest.fixture(scope="module")
def G():
    return load_golden("gms_dll")


def _mask(G, name, tag):
    n = int(G[name + "/n"])
    return np.unpackbits(G["%s/mask_%s" % (name, tag)])[:n].astype(bool)


def _full(mask, n):
    """matchesGMS cannot tell an empty mask from an all-false one: compare as length-n masks"""
    return mask if len(mask) == n else np.zeros(n, bool)


def _check_case(G, name, c, gms_fn, factor=6.0, flags=C.FLAGS):
    n = len(c["q"])
    assert n == int(G[name + "/n"])
    for tag, rot, sc in flags:
        r = gms_fn(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], rot, sc, factor)
        exp = _mask(G, name, tag)
        assert np.array_equal(_full(r["mask"], n), exp), (name, tag, int(r["n_inliers"]), int(exp.sum()))
        assert r["n_inliers"] == int(exp.sum())


def _check_hyp(G, name, c, oracle_mod, factor=6.0):
    r = oracle_mod.gms(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], True, True, factor, want_all_masks=True)
    assert r["hyp_counts"].tolist() == G[name + "/hyp_counts"].tolist(), name
    assert C.sha(r["all_masks"]) == str(G[name + "/hyp_sha"]), name


# ------------------------------------------------------------------------------------------------ CPU: oracle == DLL
def test_oracle_tables_equal_dll(G, oracle_mod):
    rot, sc = oracle_mod.gms_tables()
    assert np.array_equal(rot, G["rot"])
    assert np.array_equal(sc.view(np.uint64), G["scale"].view(np.uint64))          # bit for bit, incl. the two sqrt slots
    assert [oracle_mod.gms_right_grid(s) for s in range(5)] == C.RIGHT_GRIDS


def test_oracle_grid_index_equals_dll_on_cell_edges(G, oracle_mod):
    """every f32 within 3 ulp of a cell edge / half-cell edge of every grid, as x and as y"""
    p = C.grid_edge_points()
    for t in (1, 2, 3, 4):
        assert np.array_equal(oracle_mod.gms_grid_left(p, t), G["grid_edge_left%d" % t].astype(np.int32)), t
    for w in C.RIGHT_GRIDS:
        assert np.array_equal(oracle_mod.gms_grid_right(p, w, w), G["grid_edge_right%d" % w].astype(np.int32)), w


def test_oracle_grid_index_equals_dll_on_a_million_points(G, oracle_mod):
    p = C.grid_random_points()
    parts = [oracle_mod.gms_grid_left(p, t) for t in (1, 2, 3, 4)] + [oracle_mod.gms_grid_right(p, w, w) for w in C.RIGHT_GRIDS]
    assert C.sha(*parts) == str(G["grid_rand_sha"])


@pytest.mark.parametrize("name", C.REAL)
def test_oracle_equals_dll_real_pairs(G, oracle_mod, name):
    """ORB-10k on the reference's own images (BASELINE configs[0]) — masks for all four flag combinations and all
    40 per-hypothesis masks"""
    c = C.real_case(name)
    _check_case(G, name, c, oracle_mod.gms)
    _check_hyp(G, name, c, oracle_mod)


def test_oracle_equals_dll_sift_pair(G, oracle_mod):
    """the reference's literal pipeline: SIFT + L2 matches + matchGMS(true, true) (FeatureMatchUtil.cpp:10,66-69)"""
    c = C.sift_case()
    _check_case(G, "sift_view01_1500", c, oracle_mod.gms)
    _check_hyp(G, "sift_view01_1500", c, oracle_mod)


@pytest.mark.parametrize("cfg", ["cfg2_640x480_10k", "cfg3_1080p_50k_rs"])
def test_oracle_equals_dll_baseline_configs(G, oracle_mod, cfg):
    c = C.synth_case(cfg, lambda a, b: oracle_mod.bf_hamming(a, b)[0])
    assert C.sha(c["t"]) == str(G[cfg + "/t_sha"])
    _check_case(G, cfg, c, oracle_mod.gms)
    _check_hyp(G, cfg, c, oracle_mod)


def test_oracle_equals_dll_edge_pixels_and_subset(G, oracle_mod):
    for name, c in (("edge_pixels", C.edge_pixels_case()), ("subset", C.subset_case())):
        _check_case(G, name, c, oracle_mod.gms)
        _check_hyp(G, name, c, oracle_mod)


def test_oracle_equals_dll_threshold_factors(G, oracle_mod):
    for k, f in enumerate(C.FACTORS):
        _check_case(G, "view01_2k_f%d" % k, C.real_case("view01_2k"), oracle_mod.gms, f)
        _check_case(G, "edge_pixels_f%d" % k, C.edge_pixels_case(), oracle_mod.gms, f)
        if k % 3 == 0:
            _check_hyp(G, "view01_2k_f%d" % k, C.real_case("view01_2k"), oracle_mod, f)


def test_oracle_equals_dll_stress(G, oracle_mod):
    for i in range(C.N_STRESS):
        c = C.stress_case(i)
        _check_case(G, "stress%d" % i, c, oracle_mod.gms, c["factor"])
        if i % 4 == 0:
            _check_hyp(G, "stress%d" % i, c, oracle_mod, c["factor"])


def test_oracle_equals_dll_microcases(G, oracle_mod):
    for name, c in C.micro_inputs():
        tag = "%d%d" % (c["rot"], c["sc"])
        _check_case(G, "micro/" + name, c, oracle_mod.gms, c["factor"], [(tag, c["rot"], c["sc"])])


# ------------------------------------------------------------------------------- CPU, build container: live DLL
def _dll():
    from oracle import dllref

    if not dllref.available():
        pytest.skip("reference DLL not on this machine (build container only)")
    return dllref


def test_live_dll_matches_committed_golden(G):
    """the committed file really is what the DLL answers (guards against a stale golden)"""
    dllref = _dll()
    rot, sc = dllref.tables()
    assert np.array_equal(rot, G["rot"]) and np.array_equal(sc, G["scale"])
    c = C.real_case("view01_2k")
    for tag, r, s in C.FLAGS:
        out = dllref.match_gms(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], r, s, 6.0, tag_positions=True)
        m = np.zeros(len(c["q"]), bool)
        m[out["imgIdx"]] = True
        assert np.array_equal(m, _mask(G, "view01_2k", tag))


def test_live_dll_vs_oracle_random(oracle_mod):
    """fresh random cases (not the committed seeds): DLL export == oracle, all flags, all 40 hypotheses"""
    dllref = _dll()
    for i in range(1000, 1120):
        c = C.stress_case(i)
        n = len(c["q"])
        for tag, r, s in C.FLAGS:
            out = dllref.match_gms(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], r, s, c["factor"], tag_positions=True)
            m = np.zeros(n, bool)
            m[out["imgIdx"]] = True
            o = oracle_mod.gms(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], r, s, c["factor"])
            assert np.array_equal(_full(o["mask"], n), m), (i, tag)
        counts, masks = dllref.hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], c["factor"], True)
        o = oracle_mod.gms(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], True, True, c["factor"], want_all_masks=True)
        assert np.array_equal(counts, o["hyp_counts"]) and np.array_equal(masks, o["all_masks"]), i


def test_live_dll_grid_index_vs_oracle_fresh_points(oracle_mod):
    dllref = _dll()
    rng = np.random.default_rng(99)
    p = rng.random((1 << 21, 2), dtype=np.float32)
    for t in (1, 2, 3, 4):
        assert np.array_equal(dllref.grid_left(p, t), oracle_mod.gms_grid_left(p, t))
    for w in C.RIGHT_GRIDS:
        assert np.array_equal(dllref.grid_right(p, w, w), oracle_mod.gms_grid_right(p, w, w))


# ------------------------------------------------------------------------------------------------ GPU: CUDA == DLL
@pytest.mark.gpu
@pytest.mark.parametrize("name", C.REAL)
def test_gpu_equals_dll_real_pairs(G, ctx, name):
    c = C.real_case(name)
    _check_case(G, name, c, ctx.gms)
    assert ctx.gms_hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"]).tolist() == G[name + "/hyp_counts"].tolist()


@pytest.mark.gpu
def test_gpu_equals_dll_sift_pair(G, ctx):
    c = C.sift_case()
    _check_case(G, "sift_view01_1500", c, ctx.gms)
    assert ctx.gms_hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"]).tolist() == G["sift_view01_1500/hyp_counts"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["cfg2_640x480_10k", "cfg3_1080p_50k_rs"])
def test_gpu_equals_dll_baseline_configs(G, ctx, cfg):
    """BASELINE configs 2 and 3 at full size: CUDA BF-Hamming feeds CUDA GMS; the result equals the DLL's"""
    c = C.synth_case(cfg, lambda a, b: ctx.bf_hamming(a, b)[0])
    assert C.sha(c["t"]) == str(G[cfg + "/t_sha"])
    _check_case(G, cfg, c, ctx.gms)
    assert ctx.gms_hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"]).tolist() == G[cfg + "/hyp_counts"].tolist()


@pytest.mark.gpu
def test_gpu_equals_dll_edge_pixels_subset_factors(G, ctx):
    for name, c in (("edge_pixels", C.edge_pixels_case()), ("subset", C.subset_case())):
        _check_case(G, name, c, ctx.gms)
        assert ctx.gms_hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"]).tolist() == G[name + "/hyp_counts"].tolist()
    for k, f in enumerate(C.FACTORS):
        _check_case(G, "view01_2k_f%d" % k, C.real_case("view01_2k"), ctx.gms, f)
        _check_case(G, "edge_pixels_f%d" % k, C.edge_pixels_case(), ctx.gms, f)


@pytest.mark.gpu
def test_gpu_equals_dll_stress_and_microcases(G, ctx):
    for i in range(C.N_STRESS):
        c = C.stress_case(i)
        _check_case(G, "stress%d" % i, c, ctx.gms, c["factor"])
        if i % 4 == 0:
            h = ctx.gms_hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"], c["factor"])
            assert h.tolist() == G["stress%d/hyp_counts" % i].tolist(), i
    for name, c in C.micro_inputs():
        tag = "%d%d" % (c["rot"], c["sc"])
        _check_case(G, "micro/" + name, c, ctx.gms, c["factor"], [(tag, c["rot"], c["sc"])])


@pytest.mark.gpu
def test_gpu_dense_path_equals_dll(G, ctx):
    """the global-memory histogram path (what pairs with >= 65536 matches take) forced on: same answers"""
    from sfm_gms_b200 import api

    ctx.set_option(api.OPT_GMS_DENSE, 1)
    try:
        for name in ("view01_2k", "bun12_rot180_3k"):
            c = C.real_case(name)
            _check_case(G, name, c, ctx.gms)
            assert ctx.gms_hypotheses(c["size1"], c["size2"], c["kp1"], c["kp2"], c["q"], c["t"]).tolist() == G[name + "/hyp_counts"].tolist()
        _check_case(G, "edge_pixels", C.edge_pixels_case(), ctx.gms)
    finally:
        ctx.set_option(api.OPT_GMS_DENSE, 0)
