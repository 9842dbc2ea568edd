"""GPU tests (-m gpu) of the pair-list runner: chunked walks (bounded device memory), compacted outputs
(sfmgms_match_pairs_compact = every pair's matchesGMS vector + the coordinate gather of SfMUtil.cpp:25-35), host and
device output locations, capacity errors, and the all-pairs workload of BASELINE config 5 at a reduced size.
Everything is compared bit-exactly with the oracle (itself pinned to cv2 for stage 1 and to the reference DLL for
stage 2)."""
import numpy as np
import pytest

from test_gpu_parity import _ragged_set

pytestmark = pytest.mark.gpu

SIZES = [3000, 2500, 0, 1777, 4096, 1, 2222]
PAIRS = np.array([(0, 1), (1, 0), (0, 3), (4, 0), (2, 1), (1, 2), (5, 4), (3, 3), (4, 5), (6, 0), (0, 6), (6, 4), (1, 3), (3, 1)],
                 np.int32)


def _oracle_compact(oracle_mod, off, desc, kp, wh, pairs, rot, sc):
    """per pair: (query idx, train idx, dist, pts1, pts2) of the inliers in match order, n_inliers, best_hyp"""
    out = []
    for a, b in pairs:
        d1, d2 = desc[off[a]:off[a + 1]], desc[off[b]:off[b + 1]]
        k1, k2 = kp[off[a]:off[a + 1]], kp[off[b]:off[b + 1]]
        oi, od = oracle_mod.bf_hamming(d1, d2)
        if len(d2) == 0 or len(d1) == 0:
            out.append((np.zeros(0, int), np.zeros(0, int), np.zeros(0), k1[:0], k2[:0], 0, None))
            continue
        o = oracle_mod.gms(wh[a], wh[b], k1, k2, np.arange(len(oi), dtype=np.int32), oi, rot, sc)
        keep = np.nonzero(o["mask"])[0] if len(o["mask"]) else np.zeros(0, int)
        out.append((keep, oi[keep], od[keep], k1[keep], k2[oi[keep]], o["n_inliers"], o["best_hyp"]))
    return out


def _check_compact(res, exp):
    off = res["offsets"]
    assert off[0] == 0 and off[-1] == res["n_total"] == sum(len(e[0]) for e in exp)
    for p, (q, t, d, p1, p2, n, bh) in enumerate(exp):
        sl = slice(off[p], off[p + 1])
        assert off[p + 1] - off[p] == len(q) and res["n_inliers"][p] == n, p
        if bh is not None:
            assert res["best_hyp"][p] == bh, p
        m = res["matches"][sl]
        assert np.array_equal(m["queryIdx"], q) and np.array_equal(m["trainIdx"], t), p
        assert np.array_equal(m["distance"], d.astype(np.float32)) and not m["imgIdx"].any(), p
        assert np.array_equal(res["pts1"][sl], p1) and np.array_equal(res["pts2"][sl], p2), p


@pytest.mark.parametrize("chunk_rows", [1, 3000, 9000, 1 << 22])
@pytest.mark.parametrize("rot,sc", [(0, 0), (1, 1)])
def test_compact_ragged_vs_oracle(ctx, oracle_mod, rot, sc, chunk_rows):
    """chunk_rows = 1: every pair its own chunk; 3000 / 9000: ragged chunk borders, empty images inside chunks;
    2^22: one chunk.  Same answer every time."""
    from sfm_gms_b200 import api

    rng = np.random.default_rng(5)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    exp = _oracle_compact(oracle_mod, off, desc, kp, wh, PAIRS, rot, sc)
    ctx.set_option(api.OPT_CHUNK_ROWS, chunk_rows)
    try:
        res = ctx.match_pairs_compact(PAIRS, rot, sc)
        full = ctx.match_pairs(PAIRS, rot, sc)
    finally:
        ctx.set_option(api.OPT_CHUNK_ROWS, 4 << 20)
    _check_compact(res, exp)
    # the uncompacted outputs of a chunked walk agree with the compacted ones
    mo = full["offsets"]
    for p in range(len(PAIRS)):
        m = full["mask"][mo[p]:mo[p + 1]].astype(bool)
        assert np.array_equal(np.nonzero(m)[0], exp[p][0]), p
        assert np.array_equal(full["train_idx"][mo[p]:mo[p + 1]][m], exp[p][1]), p
    assert np.array_equal(full["n_inliers"], res["n_inliers"])


@pytest.mark.parametrize("chunk_rows", [3000, 1 << 22])
def test_compact_index_pair_records(ctx, oracle_mod, chunk_rows):
    """SFMGMS_OPT_COMPACT_RECORD = 1: `matches` holds 8-byte {queryIdx, trainIdx} rows (what SfMUtil.cpp:25-35 reads of
    matchesGMS); offsets, counts and coordinates as with cv::DMatch records."""
    from sfm_gms_b200 import api

    rng = np.random.default_rng(8)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    exp = _oracle_compact(oracle_mod, off, desc, kp, wh, PAIRS, 0, 0)
    ctx.set_option(api.OPT_CHUNK_ROWS, chunk_rows)
    try:
        r = ctx.match_pairs_compact(PAIRS, index_pairs=True)
        d = ctx.match_pairs_compact(PAIRS)
    finally:
        ctx.set_option(api.OPT_CHUNK_ROWS, 4 << 20)
    assert r["matches"].dtype == np.int32 and r["matches"].shape == (r["n_total"], 2)
    assert np.array_equal(r["matches"][:, 0], np.concatenate([e[0] for e in exp]))
    assert np.array_equal(r["matches"][:, 1], np.concatenate([e[1] for e in exp]))
    assert np.array_equal(r["offsets"], d["offsets"]) and np.array_equal(r["pts1"], d["pts1"]) and np.array_equal(r["pts2"], d["pts2"])
    _check_compact(d, exp)                                                # the option does not stick
    with pytest.raises(api.SfmGmsError):
        ctx.set_option(api.OPT_COMPACT_RECORD, 2)


def test_compact_partial_outputs_and_capacity(ctx, oracle_mod):
    from sfm_gms_b200 import api

    rng = np.random.default_rng(6)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    exp = _oracle_compact(oracle_mod, off, desc, kp, wh, PAIRS, 0, 0)
    total = sum(len(e[0]) for e in exp)
    r = ctx.match_pairs_compact(PAIRS, want_points=False)
    assert "pts1" not in r and r["n_total"] == total and np.array_equal(r["matches"]["queryIdx"], np.concatenate([e[0] for e in exp]))
    r = ctx.match_pairs_compact(PAIRS, want_matches=False)
    assert np.array_equal(r["pts2"], np.concatenate([e[4] for e in exp]))
    r = ctx.match_pairs_compact(PAIRS, capacity=total)                    # exactly enough
    assert r["n_total"] == total
    for chunk_rows in (1 << 22, 5000):
        ctx.set_option(api.OPT_CHUNK_ROWS, chunk_rows)
        try:
            with pytest.raises(api.SfmGmsError) as e:
                ctx.match_pairs_compact(PAIRS, capacity=total - 1)
            assert e.value.code == 7 and str(total) in str(e.value)       # SFMGMS_ERR_CAPACITY names the needed size
        finally:
            ctx.set_option(api.OPT_CHUNK_ROWS, 4 << 20)
    r = ctx.match_pairs_compact(np.zeros((0, 2), np.int32))
    assert r["n_total"] == 0 and len(r["offsets"]) == 1


@pytest.mark.parametrize("chunk_rows", [7000, 1 << 22])
def test_compact_device_outputs(ctx, oracle_mod, chunk_rows):
    """out_location = SFMGMS_DEVICE: results stay in caller-owned device memory (one running offset across chunks)"""
    torch = pytest.importorskip("torch")
    from sfm_gms_b200 import api

    rng = np.random.default_rng(7)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    exp = _oracle_compact(oracle_mod, off, desc, kp, wh, PAIRS, 1, 0)
    total = sum(len(e[0]) for e in exp)
    dev = torch.device("cuda", 0)
    n = len(PAIRS)
    ninl = torch.zeros(n, dtype=torch.int32, device=dev)
    bh = torch.zeros(n, dtype=torch.int32, device=dev)
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    m = torch.zeros(total * 4 + 4, dtype=torch.int32, device=dev)
    p1 = torch.zeros(total * 2 + 2, dtype=torch.float32, device=dev)
    p2 = torch.zeros(total * 2 + 2, dtype=torch.float32, device=dev)
    ctx.set_option(api.OPT_CHUNK_ROWS, chunk_rows)
    try:
        tot = ctx.match_pairs_compact_raw(np.ascontiguousarray(PAIRS), 1, 0, 6.0, api.SFMGMS_DEVICE, total, ninl.data_ptr(),
                                          bh.data_ptr(), offs.data_ptr(), m.data_ptr(), p1.data_ptr(), p2.data_ptr())
    finally:
        ctx.set_option(api.OPT_CHUNK_ROWS, 4 << 20)
    torch.cuda.synchronize()
    res = dict(n_total=tot, offsets=offs.cpu().numpy(), n_inliers=ninl.cpu().numpy(), best_hyp=bh.cpu().numpy(),
               matches=m.cpu().numpy()[: total * 4].view(api.DMATCH_DT), pts1=p1.cpu().numpy()[: total * 2].reshape(-1, 2),
               pts2=p2.cpu().numpy()[: total * 2].reshape(-1, 2))
    _check_compact(res, exp)


def test_inlier_points_is_a_view_of_the_compact_output(ctx):
    rng = np.random.default_rng(8)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    res = ctx.match_pairs_compact(PAIRS)
    co = res["offsets"]
    for p in range(len(PAIRS)):
        p1, p2, n = ctx.inlier_points(p, 5000)
        assert n == co[p + 1] - co[p]
        assert np.array_equal(p1, res["pts1"][co[p]:co[p + 1]]) and np.array_equal(p2, res["pts2"][co[p]:co[p + 1]])


def test_all_pairs_sequence_bounded_memory(oracle_mod):
    """BASELINE config 5 in miniature: all pairs over a 24-image sequence (10k keypoints each, 276 pairs) walked in
    chunks of 2^20 rows.  Sampled pairs equal the oracle; library-owned device memory stays far below the list size."""
    import sfm_gms_b200 as sg
    from sfm_gms_b200 import api, synth

    s = synth.make_sequence(24)
    pairs = synth.all_pairs(24)
    ctx = sg.Context(0)                            # a fresh context: its device memory is this test's alone
    ctx.set_images(s["offsets"], s["desc"], s["kp"], s["sizes"])
    ctx.set_option(api.OPT_CHUNK_ROWS, 1 << 19)
    res = ctx.match_pairs_compact(pairs, capacity=len(pairs) * 10_000)
    used = ctx.device_bytes
    ctx.close()
    # 276 pairs x 10k rows: per-list buffers would hold 2.76 M rows (x ~70 B); chunks of 2^19 rows keep it ~5x smaller
    assert used < 300 << 20, used
    rng = np.random.default_rng(9)
    pick = rng.choice(len(pairs), 6, replace=False)
    exp = _oracle_compact(oracle_mod, s["offsets"], s["desc"], s["kp"], s["sizes"], pairs[pick], 0, 0)
    co = res["offsets"]
    for (q, t, d, p1, p2, n, bh), p in zip(exp, pick):
        m = res["matches"][co[p]:co[p + 1]]
        assert np.array_equal(m["queryIdx"], q) and np.array_equal(m["trainIdx"], t) and res["n_inliers"][p] == n
        assert np.array_equal(res["pts1"][co[p]:co[p + 1]], p1) and np.array_equal(res["pts2"][co[p]:co[p + 1]], p2)
    # neighbours in the sequence share most landmarks, distant images share none
    near = res["n_inliers"][(pairs[:, 1] - pairs[:, 0]) == 1]
    assert near.min() > 2000


@pytest.mark.parametrize("chunk_rows", [6000, 1 << 22])
def test_async_variants_equal_sync(ctx, chunk_rows):
    """sfmgms_match_pairs_async / _compact_async + sfmgms_wait: device outputs, the call returns before the work is done;
    results equal the synchronous calls (single- and multi-chunk lists)"""
    torch = pytest.importorskip("torch")
    from sfm_gms_b200 import api

    rng = np.random.default_rng(11)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    ref = ctx.match_pairs(PAIRS, 1, 1)
    refc = ctx.match_pairs_compact(PAIRS, 1, 1)
    dev = torch.device("cuda", 0)
    n, rows, total = len(PAIRS), int(ref["offsets"][-1]), int(refc["n_total"])
    i32 = lambda k: torch.zeros(k, dtype=torch.int32, device=dev)  # noqa: E731
    ninl, bh, ml, ti, di = i32(n), i32(n), i32(n), i32(rows), i32(rows)
    mk = torch.zeros(rows, dtype=torch.uint8, device=dev)
    pr = np.ascontiguousarray(PAIRS)
    ctx.set_option(api.OPT_CHUNK_ROWS, chunk_rows)
    try:
        ctx.match_pairs_async_raw(pr, 1, 1, 6.0, ninl.data_ptr(), bh.data_ptr(), ml.data_ptr(), ti.data_ptr(), di.data_ptr(), mk.data_ptr())
        with pytest.raises(api.SfmGmsError):           # the context is busy until wait()
            ctx.match_pairs(PAIRS)
        assert ctx.wait() == total
        for a, b in ((ninl, "n_inliers"), (bh, "best_hyp"), (ml, "mask_len"), (ti, "train_idx"), (di, "dist"), (mk, "mask")):
            assert np.array_equal(a.cpu().numpy(), ref[b]), b
        offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        m = i32(total * 4 + 4)
        p1 = torch.zeros(total * 2 + 2, dtype=torch.float32, device=dev)
        ninl.zero_()
        ctx.match_pairs_compact_async_raw(pr, 1, 1, 6.0, total, ninl.data_ptr(), bh.data_ptr(), offs.data_ptr(), m.data_ptr(), p1.data_ptr())
        assert ctx.wait() == total
        assert np.array_equal(offs.cpu().numpy(), refc["offsets"]) and np.array_equal(ninl.cpu().numpy(), refc["n_inliers"])
        assert np.array_equal(m.cpu().numpy()[: total * 4].view(api.DMATCH_DT), refc["matches"])
        assert np.array_equal(p1.cpu().numpy()[: total * 2].reshape(-1, 2), refc["pts1"])
        # too small a buffer is reported by wait()
        ctx.match_pairs_compact_async_raw(pr, 1, 1, 6.0, total - 1, ninl.data_ptr(), bh.data_ptr(), offs.data_ptr(), m.data_ptr())
        with pytest.raises(api.SfmGmsError) as e:
            ctx.wait()
        assert e.value.code == 7
        assert ctx.wait() == 0                           # nothing pending any more
    finally:
        ctx.set_option(api.OPT_CHUNK_ROWS, 4 << 20)


@pytest.mark.gpu
def test_host_alloc_buffers_round_trip(ctx):
    """sfmgms_host_alloc memory as the image set and the result tables: same results as ordinary numpy arrays, and the
    block is page-locked (cudaHostAlloc) -- torch reports it as pinned."""
    torch = pytest.importorskip("torch")
    from sfm_gms_b200 import api

    rng = np.random.default_rng(5)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    ctx.set_images(off, desc, kp, wh)
    ref = ctx.match_pairs(PAIRS, 1, 0)
    h_desc, h_kp = api.host_array(desc), api.host_array(kp)
    assert h_desc.flags.c_contiguous and np.array_equal(h_desc, desc)
    assert torch.from_numpy(h_desc).is_pinned() and torch.from_numpy(h_kp).is_pinned()
    ctx.set_images(off, h_desc, h_kp, wh)
    got = ctx.match_pairs(PAIRS, 1, 0)
    for k in ("n_inliers", "best_hyp", "train_idx", "dist", "mask"):
        assert np.array_equal(got[k], ref[k]), k
    n, rows = len(PAIRS), int(ref["offsets"][-1])
    outs = [api.host_zeros(n, np.int32) for _ in range(3)] + [api.host_zeros(rows, np.int32) for _ in range(2)] + \
           [api.host_zeros(rows, np.uint8)]
    ctx.match_image_set_raw(off, h_desc.ctypes.data, h_kp.ctypes.data, wh, np.ascontiguousarray(PAIRS), 1, 0, 6.0,
                            *[o.ctypes.data for o in outs])
    for o, k in zip(outs, ("n_inliers", "best_hyp", "mask_len", "train_idx", "dist", "mask")):
        assert np.array_equal(o, ref[k]), k
    z = api.host_empty((0, 32), np.uint8)
    assert z.shape == (0, 32)
    del h_desc, h_kp, outs                                 # blocks are released with their last view
