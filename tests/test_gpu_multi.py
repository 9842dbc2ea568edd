"""GPU tests (-m gpu) of the multi-GPU C ABI (sfmgms_multi_*): one process, one host thread per GPU, the set broadcast
once over NCCL, the pair list sharded.  Results must equal the single-GPU calls bit for bit.  With one visible GPU the
same code runs as a 1-GPU "multi" (no NCCL); with >= 2 (gpurun --gpus 2) the broadcast and the sharding are exercised."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from test_gpu_parity import _ragged_set

pytestmark = pytest.mark.gpu

SIZES = [3000, 2500, 0, 1777, 4096, 1, 2222, 3100]


def _gpu_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("n_dev", [1, 2, 4])
@pytest.mark.parametrize("rot,sc", [(0, 0), (1, 1)])
def test_multi_equals_single(ctx, n_dev, rot, sc):
    import sfm_gms_b200 as sg
    from sfm_gms_b200 import api

    if n_dev > _gpu_count():
        pytest.skip("needs %d GPUs" % n_dev)
    rng = np.random.default_rng(21)
    off, desc, kp, wh = _ragged_set(rng, SIZES)
    n = len(SIZES)
    pairs = np.array([(i, j) for i in range(n) for j in range(n) if i != j and (i + j) % 3], np.int32)
    ctx.set_images(off, desc, kp, wh)
    ref = ctx.match_pairs(pairs, rot, sc)
    refc = ctx.match_pairs_compact(pairs, rot, sc)
    m = api.MultiContext(list(range(n_dev)))
    try:
        assert m.n_devices == n_dev
        m.set_option(api.OPT_CHUNK_ROWS, 9000)          # several chunks per GPU: chunks of different GPUs interleave
        m.set_images(off, desc, kp, wh)
        full = m.match_pairs(pairs, rot, sc)
        for k in ["n_inliers", "best_hyp", "mask_len", "offsets", "train_idx", "dist", "mask"]:
            assert np.array_equal(full[k], ref[k]), k
        c = m.match_pairs_compact(pairs, rot, sc)
        assert c["n_total"] == refc["n_total"] and np.array_equal(c["n_inliers"], refc["n_inliers"])
        seen = np.zeros(c["n_total"], bool)
        for p in range(len(pairs)):
            a = slice(c["begin"][p], c["begin"][p] + c["n_inliers"][p])
            b = slice(refc["offsets"][p], refc["offsets"][p + 1])
            assert np.array_equal(c["matches"][a], refc["matches"][b]), p
            assert np.array_equal(c["pts1"][a], refc["pts1"][b]) and np.array_equal(c["pts2"][a], refc["pts2"][b]), p
            assert not seen[a].any()
            seen[a] = True
        assert seen.all()                               # the rows of all pairs tile the output exactly
        with pytest.raises(api.SfmGmsError) as e:
            m.match_pairs_compact(pairs, rot, sc, capacity=max(0, refc["n_total"] - 1))
        assert e.value.code == 7
    finally:
        m.close()


def test_multi_argument_errors():
    from sfm_gms_b200 import api

    with pytest.raises(api.SfmGmsError):
        api.MultiContext([0, 0])                        # a device listed twice
    with pytest.raises(api.SfmGmsError):
        api.MultiContext([_gpu_count() + 3])
    m = api.MultiContext([0])
    try:
        with pytest.raises(api.SfmGmsError):
            m.match_pairs(np.array([[0, 1]], np.int32))  # no image set yet
    finally:
        m.close()


@pytest.mark.parametrize("rot,sc", [(0, 0), (1, 1)])
def test_cxx_multi_gpu_demo(rot, sc):
    """the C++ host (sfmgms::MultiGpuMatcher in sfmgms.hpp) on all visible GPUs against sfmgms::matchPairs on one"""
    exe = os.path.join(ROOT, "sfm_gms_b200", "cxx", "demo_multi")
    if not os.path.exists(exe):
        from sfm_gms_b200 import build_cxx

        build_cxx.build()
    out = subprocess.run([exe, "0", "10", "3000", str(rot), str(sc)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["mismatching_pairs"] == 0 and r["gpus"] == _gpu_count() and r["pairs"] == 45 and r["inliers"] > 1000
