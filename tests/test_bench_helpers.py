"""CPU tests of bench.py's roofline bookkeeping (no GPU): the byte / operation models the JSON line reports must be the
ones SURVEY §8(d) and DESIGN.md §4 state."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_gms_stage_figures_match_the_survey():
    b = _bench()
    # figure (i) 0.25 MB per cfg2 pair; figure (ii) ~5.6 MB (cfg2, H=1), ~67 MB (cfg3, 40 hypotheses), ~14.7 MB (cfg4, H=1)
    r2 = b.gms_stage_roofline(0.15, 256, 10_000, 1, 1)
    assert r2["compulsory_bytes_per_pair"] == 250_000 and abs(r2["reference_algorithm_bytes_per_pair"] - 5.69e6) < 1e4
    assert r2["hypotheses"] == 1 and r2["bound"] == "hbm"
    r3 = b.gms_stage_roofline(0.55, 32, 50_000, 5, 8)
    assert r3["hypotheses"] == 40 and abs(r3["reference_algorithm_bytes_per_pair"] - 67.5e6) < 1e5
    r4 = b.gms_stage_roofline(0.19, 4, 200_000, 1, 1)
    assert abs(r4["reference_algorithm_bytes_per_pair"] - 16.52e6) < 1e5
    # fractions follow from bytes / time / peak
    assert abs(r2["frac_compulsory"] - 256 * 250_000 / 0.15e-3 / 1e9 / r2["peak"]) < 1e-12


def test_hamming_roofline_uses_the_measured_unit_peak():
    b = _bench()
    dists = 256.0 * 10_000 * 10_000
    r = b.hamming_roofline("fp4", dists, 2.0, None)
    assert r["bound"] == "tensor" and r["kernel"] == "hamming_fp4_kernel" and r["unit"] == "TOP/s"
    assert abs(r["achieved"] - 2 * 256 * dists / 2.0e-3 / 1e12) < 1e-6
    unit = b.load_json("profiles", "r2_unit_peaks.json")
    assert unit.get("mxf4_tops") and r["peak"] == unit["mxf4_tops"] and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    p = b.hamming_roofline("popc", dists, 50.0, None)
    assert p["bound"] == "int_popc" and abs(p["achieved"] - 8 * dists / 50e-3 / 1e9) < 1e-6


def test_kernel_rooflines_models():
    b = _bench()
    kt = {"hamming_fp4": (5.4, 3), "unpack_fp4": (0.234, 3), "gms_vote2": (0.246, 3), "gms_assign_cnt": (0.129, 3), "mystery": (0.01, 3)}
    rows = {e["kernel"]: e for e in b.kernel_rooflines(kt, 3, "fp4", 256, 10_000, 1)}
    assert list(rows)[0] == "hamming_fp4"                      # sorted by time
    u = rows["unpack_fp4"]                                       # train images only: 160 B per row of ONE image per pair
    assert u["bound"] == "hbm" and u["algorithmic_bytes_per_step"] == 256 * 10_000 * 160
    v = rows["gms_vote2"]
    assert v["bound"] == "smem_atomic" and v["algorithmic_votes_per_step"] == 256 * 10_000 * 4
    assert "frac" not in rows["mystery"] and rows["mystery"]["launches_per_step"] == 1.0
