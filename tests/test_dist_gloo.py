"""CPU test of the N>1 path: world_size-2 gloo, the same sharding / broadcast / gather code bench.py and users run
with NCCL.  The per-pair compute is stubbed with the CPU oracle (test infrastructure) so the plumbing can be checked
without a GPU: sharded results must equal the single-process results pair by pair."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
import oracle
from sfm_gms_b200 import synth, dist as sd

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mode = sys.argv[2]
image_set = synth.make_sequence(5, n_kp=600, w=320, h=240) if rank == 0 else None
pairs = synth.all_pairs(5)
dev = torch.device("cpu")
s = sd.broadcast_image_set(image_set, 0, dev)
assert s["desc"].shape == (3000, 32) and s["kp"].dtype == torch.float32

def compute(iset, local_pairs):
    off, desc, kp = iset["offsets"], iset["desc"].numpy(), iset["kp"].numpy()
    ninl, bh = [], []
    for a, b in local_pairs:
        d1, d2 = desc[off[a]:off[a+1]], desc[off[b]:off[b+1]]
        idx, _ = oracle.bf_hamming(d1, d2)
        r = oracle.gms(iset["sizes"][a], iset["sizes"][b], kp[off[a]:off[a+1]], kp[off[b]:off[b+1]],
                       np.arange(len(idx), dtype=np.int32), idx, True, False)
        ninl.append(r["n_inliers"]); bh.append(r["best_hyp"])
    return dict(n_inliers=ninl, best_hyp=bh)

idx, res, summ = sd.match_pairs_sharded(compute, s, pairs, shard_mode=mode)
assert len(idx) in (5, 5) and len(res["n_inliers"]) == len(idx)
if rank == 0:
    full = compute(s, pairs)                     # single-process answer on the same broadcast set
    assert summ.shape == (10, 2)
    assert summ[:, 0].tolist() == list(full["n_inliers"]) and summ[:, 1].tolist() == list(full["best_hyp"])
    assert (summ[:, 0] > 50).all()
    print("OK", mode, summ[:, 0].tolist())
else:
    assert summ is None
dist.barrier()
dist.destroy_process_group()
'''


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["contiguous", "strided"])
def test_sharded_matching_world2_gloo(tmp_path, mode):
    pytest.importorskip("torch")
    import oracle

    oracle.build()
    w = tmp_path / "worker.py"
    w.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(w), ROOT, mode]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "OK " + mode in p.stdout


def test_shard_pairs_partitions_exactly():
    from sfm_gms_b200 import dist as sd, synth

    pairs = synth.all_pairs(9)
    for world in (1, 2, 3, 4, 8):
        for mode in ("contiguous", "strided"):
            seen = np.concatenate([sd.shard_pairs(pairs, r, world, mode)[0] for r in range(world)])
            assert sorted(seen.tolist()) == list(range(len(pairs)))
            sizes = [len(sd.shard_pairs(pairs, r, world, mode)[0]) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    w = synth.window_pairs(6, 2)
    assert w.tolist() == [[0, 1], [0, 2], [1, 2], [1, 3], [2, 3], [2, 4], [3, 4], [3, 5], [4, 5]]
