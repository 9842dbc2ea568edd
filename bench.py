#!/usr/bin/env python
"""bench.py — pairs/sec of the BF-Hamming + GMS hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload auto|cfg2|cfg3|cfg4|allpairs] [--impl reference]

Workloads (BASELINE.json configs; per-pair work of cfg2 and of the all-pairs job is identical: 10k x 10k, GMS defaults)
  cfg2      configs[1]: a batch of 256 DISTINCT synthetic 640x480 pairs per step and per GPU (10k keypoints x 256-bit
            descriptors per image).  The N=1 headline (`--workload auto` at N=1).  At N>1: weak scaling of
            independent batches (no collective on the data path).
  allpairs  configs[4]: all 130,816 pairs of a 512-image synthetic sequence, FIXED total work pair-sharded over the
            ranks (strong scaling): host image set -> one upload on rank 0 -> one NCCL broadcast -> every rank matches
            its contiguous shard of the (i-major) pair list -> every pair's matchesGMS vector lands in that rank's
            pinned host memory.  The N>1 headline; at N=1 it is reported beside cfg2 under "allpairs".
  cfg3/cfg4 configs[2]/[3]: 50k keypoints with the 40-hypothesis search / 200k keypoints; batches of 32 / 4 pairs.

Keys: `value` = whole-job pairs/s with the inputs resident in HBM; `e2e` = the same job through the C ABI with HOST
buffers (H2D of the inputs and D2H of the results inside the timed region); `roofline` = dominant kernel;
`roofline_kernels` = every kernel of the step (per-kernel CUDA events, separate untimed pass); `cpu_baseline` /
`--impl reference` = the reference's CPU path on a bounded sample of the same pairs (same generator, same seeds), whose
results are ALSO the parity check of the GPU arm (`parity_checked_pairs`).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image pairs/sec (ORB-10k, BF-Hamming+GMS)"
WORKLOADS = {
    "cfg2": "cfg2: synthetic 640x480 pairs, 10k kpts x 256-bit desc per image, BF-Hamming + GMS defaults",
    "cfg3": "cfg3: synthetic 1920x1080 pairs, 50k kpts x 256-bit desc per image, BF-Hamming + GMS withRotation+withScale (40 hypotheses)",
    "cfg4": "cfg4: synthetic 3840x2160 pairs, 200k kpts x 256-bit desc per image, BF-Hamming + GMS defaults",
    "allpairs": "cfg5: all-pairs matching over a 512-image synthetic sequence (130,816 pairs), 10k kpts x 256-bit desc per image, "
                "BF-Hamming + GMS defaults, pair-sharded over the GPUs",
}
# (width, height, keypoints, rotation, scale, pairs per step, make_pair kwargs, first seed)
BATCH = {
    "cfg2": (640, 480, 10_000, 0, 0, 256, {}, 2),
    "cfg3": (1920, 1080, 50_000, 1, 1, 32, dict(rot_k=2, scale=0.5, shift_frac=0.0), 3),
    "cfg4": (3840, 2160, 200_000, 0, 0, 4, {}, 4),
}
SEQ_IMAGES, SEQ_KP, SEQ_W, SEQ_H = 512, 10_000, 640, 480


def pick_workload(args, world):
    if args.workload != "auto":
        return args.workload
    return "cfg2" if world == 1 else "allpairs"


# ------------------------------------------------------------------------------------------------ inputs
def make_batch(kind, n_pairs, first_pair=0):
    """n_pairs distinct pairs of workload `kind` as one image set: images (2p, 2p+1) form pair p.  Pair p of the global
    list uses seed seed0 + p, so the reference arm (first pairs of the list) sees the same bytes."""
    from sfm_gms_b200 import synth

    w, h, n, _, _, _, kw, seed0 = BATCH[kind]
    descs, kps = [], []
    for p in range(first_pair, first_pair + n_pairs):
        d = synth.make_pair(w, h, n, seed0 + p, **kw)
        descs += [d["desc1"], d["desc2"]]
        kps += [d["kp1"], d["kp2"]]
    return dict(offsets=np.arange(2 * n_pairs + 1, dtype=np.int64) * n, desc=np.concatenate(descs), kp=np.concatenate(kps),
                sizes=np.tile(np.array([[w, h]], np.int32), (2 * n_pairs, 1)),
                pairs=np.ascontiguousarray(np.arange(2 * n_pairs, dtype=np.int32).reshape(-1, 2)))


def make_sequence(n_images):
    from sfm_gms_b200 import synth

    s = synth.make_sequence(n_images, SEQ_KP, SEQ_W, SEQ_H)
    s["pairs"] = synth.all_pairs(n_images)
    return s


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region (B200_PROFILING.md clocks line), sampled every
    few ms through NVML (the counters nvidia-smi prints; nvidia-smi's own 100 ms minimum period is too coarse)."""

    def __init__(self, gpu_index):
        self.idx, self.samples, self.stop_flag, self.t = gpu_index, [], False, None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def _run(self):
        nv = self.nv
        k, pw = 0, 0.0
        while not self.stop_flag:
            try:
                if k % 4 == 0:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), pw,
                                     nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)))
            except Exception:
                pass
            k += 1
            time.sleep(0.005)

    def stop(self):
        if self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=1)
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(s[2] & bit for s in self.samples))
        sm = [s[0] for s in self.samples]
        pw = [s[1] for s in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm), "reasons": reasons,
                "power_w_median": float(np.median(pw)) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference path
class CpuPath:
    """The reference's CPU implementation of the path: BFMatcher(NORM_HAMMING).match (OpenCV through cv2 when importable,
    else / or the C oracle port — whichever is faster on this box, decided on the first pairs) + matchGMS (C oracle port,
    single-threaded as OpenCV's GMS is; pinned to the reference DLL's own code, tests/test_gms_dll.py)."""

    def __init__(self, threads):
        import oracle

        oracle.build()
        oracle.set_num_threads(threads)
        self.oracle, self.threads = oracle, threads
        try:
            import cv2

            cv2.setNumThreads(threads)
            self.cv2, self.bf = cv2, cv2.BFMatcher(cv2.NORM_HAMMING, False)
        except Exception:
            self.cv2, self.bf = None, None
        self.use_cv2 = None
        self.t_bf = self.t_gms = 0.0
        self.t_alt = [0.0, 0.0]
        self.n = 0

    def _bf(self, d1, d2):
        o = self.oracle
        if self.bf is not None and self.use_cv2 is None and self.n < 3:      # probe both on the first pairs
            t0 = time.perf_counter()
            idx, dist = o.bf_hamming(d1, d2)
            t1 = time.perf_counter()
            m = self.bf.match(d1, d2)
            t2 = time.perf_counter()
            assert m[len(m) // 2].trainIdx == idx[len(m) // 2]
            self.t_alt[0] += t1 - t0
            self.t_alt[1] += t2 - t1
            if self.n == 2:
                self.use_cv2 = self.t_alt[1] < self.t_alt[0]
            return idx, dist, min(t1 - t0, t2 - t1)
        t0 = time.perf_counter()
        if self.use_cv2:
            m = self.bf.match(d1, d2)
            t = time.perf_counter() - t0
            idx = np.fromiter((x.trainIdx for x in m), np.int32, len(m))
            dist = np.fromiter((int(x.distance) for x in m), np.int32, len(m))
            return idx, dist, t
        idx, dist = o.bf_hamming(d1, d2)
        return idx, dist, time.perf_counter() - t0

    def pair(self, d1, d2, k1, k2, size1, size2, rot, sc):
        idx, dist, t = self._bf(d1, d2)
        self.t_bf += t
        t0 = time.perf_counter()
        g = self.oracle.gms(size1, size2, k1, k2, np.arange(len(idx), dtype=np.int32), idx, rot, sc)
        self.t_gms += time.perf_counter() - t0
        self.n += 1
        return idx, dist, g

    def describe(self):
        which = ("cv2.BFMatcher(NORM_HAMMING) %s" % self.cv2.__version__) if self.use_cv2 else "C oracle port (pthreads, popcnt)"
        return ("BF = %s on %d threads (%.1f ms/pair), GMS = C oracle port, 1 thread (%.2f ms/pair)"
                % (which, self.threads, 1e3 * self.t_bf / max(1, self.n), 1e3 * self.t_gms / max(1, self.n)))


def image_views(s, a):
    o = s["offsets"]
    return s["desc"][o[a]:o[a + 1]], s["kp"][o[a]:o[a + 1]], tuple(int(v) for v in s["sizes"][a])


def cpu_pairs(cpu, s, pairs, rot, sc):
    out = []
    for a, b in pairs:
        d1, k1, s1 = image_views(s, a)
        d2, k2, s2 = image_views(s, b)
        out.append(cpu.pair(d1, d2, k1, k2, s1, s2, rot, sc))
    return out


def sample_pairs(kind, n, seed=12345):
    """the bounded sample the CPU arms run: the first n pairs of a batch workload / n seeded random pairs of the
    all-pairs list (the GPU arm checks exactly these)"""
    if kind != "allpairs":
        return np.arange(n)
    from sfm_gms_b200 import synth

    total = len(synth.all_pairs(SEQ_IMAGES))
    return np.sort(np.random.default_rng(seed).choice(total, n, replace=False))


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path, all host threads, rank 0 only."""
    if rank != 0:
        return
    kind = pick_workload(args, world)
    threads = os.cpu_count() or 1
    n = args.ref_pairs
    if kind == "allpairs":
        s = make_sequence(SEQ_IMAGES)
        pairs = s["pairs"][sample_pairs(kind, n)]
        rot = sc = 0
    else:
        s = make_batch(kind, n)
        pairs = s["pairs"]
        rot, sc = BATCH[kind][3], BATCH[kind][4]
    cpu = CpuPath(threads)
    for _ in range(args.warmup):
        cpu_pairs(cpu, s, pairs[:1], rot, sc)
    cpu.t_bf = cpu.t_gms = 0.0
    cpu.n = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pairs(cpu, s, pairs, rot, sc)
    t_total = cpu.t_bf + cpu.t_gms            # the path itself (BF + GMS), as FeatureMatchUtil.cpp:65-71 times it
    wall = time.perf_counter() - t0
    pairs_s = n * args.steps / t_total
    line = {"impl": "reference", "metric": METRIC, "value": pairs_s, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if kind == "allpairs" else "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": WORKLOADS[kind]},
            "details": {"pairs_per_step": n, "wall_s": wall, "note": "CPU path, rank 0 only; each step = a bounded sample of the workload"},
            "cpu_baseline": {"value": pairs_s, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": "each step = %d pairs of the workload (same generator and seeds as the GPU arm); %s" % (n, cpu.describe())},
            "e2e": {"value": pairs_s, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ roofline helpers
def load_json(*path):
    try:
        return json.load(open(os.path.join(ROOT, *path)))
    except Exception:
        return {}


def hamming_roofline(kind, dists_per_step, ham_ms_per_step, clocks):
    """dominant kernel: algorithmic work = 2*256 OPs per distance evaluation on the +-1 operands (tensor kernels) or
    8 POPC.b32 per distance (popc kernel), over the CUDA-event duration of the Hamming stage on the library stream"""
    peaks = load_json("MEASURED_PEAKS.json")
    unit = load_json("profiles", "r2_unit_peaks.json")
    per_s = ham_ms_per_step * 1e-3
    if kind == "popc":
        achieved = 8 * dists_per_step / per_s / 1e9
        meas = unit.get("popc_gpopc_s")
        peak = meas or 16 * 148 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e9
        return {"bound": "int_popc", "kernel": "hamming_popc_kernel", "achieved": achieved, "peak": peak, "unit": "Gpopc/s",
                "frac": achieved / peak, "traffic": None,
                "peak_source": "profiles/r2_unit_peaks.json (POPC loop microbenchmark, this pool's B200)" if meas else
                               "nominal 16 POPC lanes/clk/SM x 148 SMs x sm_max_mhz"}
    ops = 2.0 * 256 * dists_per_step
    achieved = ops / per_s / 1e12
    mult = 4.0 if kind == "fp4" else 2.0
    key = "mxf4_tops" if kind == "fp4" else "i8_tops"
    meas = unit.get(key)
    burst = mult * peaks.get("bf16_tflops", 1660.1)
    peak = meas or burst
    roof = {"bound": "tensor", "kernel": "hamming_%s_kernel" % kind, "achieved": achieved, "peak": peak, "unit": "TOP/s",
            "frac": achieved / peak,
            "peak_source": ("profiles/r2_unit_peaks.json: tcgen05.mma %s issue loop from shared memory, no epilogue, measured on "
                            "this pool's B200 (scripts/unit_peaks.py)" % ("kind::mxf4 M128 N240 K64" if kind == "fp4" else "kind::i8")) if meas
            else "%gx measured burst bf16 TFLOP/s (MEASURED_PEAKS.json) -- no unit microbenchmark record found" % mult,
            "frac_of_%gx_burst_bf16" % mult: achieved / burst, "frac_of_nominal": achieved / (mult * 2250.0), "traffic": None,
            "avg_stage_ms": ham_ms_per_step}
    return roof


def gms_stage_roofline(gms_ms_per_step, n_pairs, n_kp, n_scales, n_rot):
    """SURVEY §8d asks for two HBM figures for the GMS passes, both over the CUDA-event duration of the whole GMS stage:
    (i) compulsory bytes per pair = 8(N1+N2) keypoints + 8N match indices + N mask bytes; (ii) the bytes the reference
    algorithm moves when its histograms live in global memory = per (scale, shift) histogram [zero + scan 4*400*G_r(s) each +
    12N assign/RMW] + 8*N*H mark reads + N.  The kernels keep the histograms in shared memory, so (ii) / time may exceed the
    HBM peak: it says how far the stage is from a global-memory implementation, (i) how far from its own floor."""
    hbm = load_json("MEASURED_PEAKS.json").get("hbm_gbs", 6532.9)
    N = float(n_kp)
    g_r = [400, 100, 196, 784, 1600][:n_scales]
    H = n_scales * n_rot
    b1 = 8 * (2 * N) + 8 * N + N
    b2 = sum(4 * (2 * 4 * 400 * g + 12 * N) for g in g_r) + 8 * N * H + N
    per_s = gms_ms_per_step * 1e-3
    return {"bound": "hbm", "stage": "gms (assign + vote + count + select)", "ms_per_step": gms_ms_per_step, "peak": hbm, "unit": "GB/s",
            "compulsory_bytes_per_pair": b1, "achieved_compulsory": n_pairs * b1 / per_s / 1e9, "frac_compulsory": n_pairs * b1 / per_s / 1e9 / hbm,
            "reference_algorithm_bytes_per_pair": b2, "achieved_reference_algorithm": n_pairs * b2 / per_s / 1e9,
            "frac_reference_algorithm": n_pairs * b2 / per_s / 1e9 / hbm, "hypotheses": H}


def kernel_rooflines(kt, steps, kind, n_pairs, n_kp, n_hyp_scales):
    """per-kernel roofline entries from the per-kernel event pass (kt: name -> (total ms, launches) over `steps` steps).
    HBM-bound kernels: algorithmic bytes / time vs the measured copy bandwidth; the shared-memory histogram kernel:
    shared-atomic operations / time vs the measured shared-atomic rate (scripts/unit_peaks.py)."""
    peaks = load_json("MEASURED_PEAKS.json")
    unit = load_json("profiles", "r2_unit_peaks.json")
    hbm = peaks.get("hbm_gbs", 6532.9)
    rows = float(n_pairs) * n_kp
    S = n_hyp_scales
    model = {   # algorithmic bytes per step: reads + writes each kernel cannot avoid
        # batches derive operands for the TRAIN images only (the kernel expands its query tiles from the packed rows)
        "unpack_fp4": ("hbm", (1 if os.environ.get("SFMGMS_FP4_PACKED_QUERIES", "1") != "0" else 2) * rows * (32 + 128)),
        "unpack_pm1": ("hbm", 2 * rows * (32 + 256)),
        "hamming_resolve": ("hbm", rows * (4 + 32 + 8 * 32 + 4)),
        "gms_assign": ("hbm", rows * (4 + 8 + 8 + 2 * (4 + S))),
        "gms_assign_cnt": ("hbm", rows * (4 + 8 + 8 + 2 * (4 + S))),
        "gms_count": ("hbm", rows * (2 * (4 + S) + 1)),
        "gms_count_scale": ("hbm", rows * (2 * (4 + S))),
        "gms_mask": ("hbm", rows * (2 * 5 + 1)),
        "gms_compact": ("hbm", rows * (1 + 4) + 0.5 * rows * (16 + 32)),
        "decode_keys": ("hbm", rows * 12),
    }
    out = []
    for name, (ms, n) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
        per_step_ms = ms / steps
        e = {"kernel": name, "ms_per_step": per_step_ms, "launches_per_step": n / steps}
        if name in model:
            b = model[name][1]
            e.update(bound="hbm", achieved=b / (per_step_ms * 1e-3) / 1e9, peak=hbm, unit="GB/s", algorithmic_bytes_per_step=b)
            e["frac"] = e["achieved"] / hbm
        elif name.startswith("gms_vote"):
            # shared-memory histogram build: 4 shifts x S scales x rows votes (plus halo duplicates, not counted)
            ops = rows * 4 * S
            e.update(bound="smem_atomic", achieved=ops / (per_step_ms * 1e-3) / 1e9, unit="Gvotes/s", algorithmic_votes_per_step=ops)
            pk = unit.get("smem_vote_gvotes") if name == "gms_vote2" else (unit.get("smem_atomic_gops", 0) / 2 or None)
            if pk:
                e.update(peak=pk, frac=e["achieved"] / pk,
                         peak_source="profiles/r2_unit_peaks.json: shared-memory vote microbenchmark (returning atomicAdd on a random "
                                     "histogram word + atomicMax on a row slot per vote, 2 CTAs x 512 threads per SM)")
        out.append(e)
    return out


# ------------------------------------------------------------------------------------------------ GPU arms
class Dist:
    def __init__(self, rank, world, local_rank):
        import torch

        self.torch, self.rank, self.world = torch, rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        if world > 1:
            import torch.distributed as dist

            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_ms(self, ms):
        t = self.torch.tensor([ms], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_i(self, v):
        t = self.torch.tensor([int(v)], device=self.dev, dtype=self.torch.int64)
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def min_i(self, v):
        t = self.torch.tensor([int(v)], device=self.dev, dtype=self.torch.int64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return int(t.item())


def timed_device(D, stream, fn, steps):
    """K steps bracketed by barrier + synchronize, CUDA events on the library's stream, max over ranks"""
    torch = D.torch
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    D.barrier()
    return D.max_ms(e0.elapsed_time(e1))


def timed_host_calls(D, fn, steps):
    """K blocking end-to-end calls: events on the idle default stream bracket them (device time == wall time here,
    both reported), max over ranks"""
    torch = D.torch
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    e1.record()
    D.barrier()
    return D.max_ms(e0.elapsed_time(e1)), D.max_ms(1e3 * wall)


def _pinned(a):
    """Host tensor in page-locked memory from the library's own allocator (sfmgms_host_alloc), holding a copy of `a`."""
    import torch
    from sfm_gms_b200 import api
    return torch.from_numpy(api.host_array(a))


def _pinned_zeros(n, dtype):
    import torch
    from sfm_gms_b200 import api
    return torch.from_numpy(api.host_zeros(int(n), dtype))


def make_ctx(args, local_rank):
    import sfm_gms_b200 as sg
    from sfm_gms_b200 import api

    ctx = sg.Context(local_rank)
    if args.kernel != "auto":
        ctx.set_option(api.OPT_HAMMING_KERNEL, {"popc": api.HAMMING_POPC, "tc": api.HAMMING_TC, "fp4": api.HAMMING_FP4}[args.kernel])
    return ctx


def run_batch_workload(args, D, kind, local_rank):
    """cfg2 / cfg3 / cfg4: P distinct pairs per step and per GPU"""
    import torch

    from sfm_gms_b200 import api

    w, h, n_kp, rot, sc, P, _, _ = BATCH[kind]
    P = args.pairs or P
    rank, world, dev = D.rank, D.world, D.dev
    s = make_batch(kind, P, first_pair=rank * P)
    total_pairs = P * world
    tot_m = P * n_kp
    h_desc = _pinned(s["desc"])
    h_kp = _pinned(s["kp"])
    d_desc, d_kp = h_desc.to(dev), h_kp.to(dev)
    ctx = make_ctx(args, local_rank)
    ctx.set_option(api.OPT_TIMING, 1)
    ctx.set_option(api.OPT_TC_OPERAND_CACHE, 0)   # a new batch every step in production: re-derive the +-1 operands every step
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    o_ninl, o_bh, o_ml = (torch.zeros(P, dtype=torch.int32, device=dev) for _ in range(3))
    o_ti, o_di = (torch.zeros(tot_m, dtype=torch.int32, device=dev) for _ in range(2))
    o_mk = torch.zeros(tot_m, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.set_images_raw(s["offsets"], d_desc.data_ptr(), d_kp.data_ptr(), s["sizes"], api.SFMGMS_DEVICE, keepalive=(d_desc, d_kp))

    def step_resident():
        ctx.match_pairs_raw(s["pairs"], rot, sc, 6.0, api.SFMGMS_DEVICE, o_ninl.data_ptr(), o_bh.data_ptr(), o_ml.data_ptr(),
                            o_ti.data_ptr(), o_di.data_ptr(), o_mk.data_ptr())

    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = ctx.kernel_launches
    ham_ms, gms_ms = [], []

    def step_resident_timed():
        step_resident()
        a, b, _ = ctx.last_timing()
        ham_ms.append(a)
        gms_ms.append(b)

    ms_total = timed_device(D, stream, step_resident_timed, args.steps)
    launches = ctx.kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    value = total_pairs * args.steps / (ms_total * 1e-3)
    res_dev = dict(n_inliers=o_ninl.cpu().numpy().copy(), train_idx=o_ti.cpu().numpy().copy(), dist=o_di.cpu().numpy().copy(),
                   mask=o_mk.cpu().numpy().copy(), mask_len=o_ml.cpu().numpy().copy())

    # per-kernel pass (untimed): one event after every kernel
    ctx.set_option(api.OPT_TIMING, 2)
    ctx.kernel_times()
    for _ in range(3):
        step_resident()
    kt = ctx.kernel_times()
    ctx.set_option(api.OPT_TIMING, 1)

    # ---- e2e: host (pinned) buffers through the C ABI, H2D + D2H inside the timed region, every step ------------
    h_out = [_pinned_zeros(P, np.int32) for _ in range(3)] + \
            [_pinned_zeros(tot_m, np.int32) for _ in range(2)] + [_pinned_zeros(tot_m, np.uint8)]
    h2d = h_desc.numel() + h_kp.numel() * 4
    d2h = 3 * P * 4 + tot_m * 9

    def call_e2e(c, outs):
        c.match_image_set_raw(s["offsets"], h_desc.data_ptr(), h_kp.data_ptr(), s["sizes"], s["pairs"], rot, sc, 6.0,
                              *[o.data_ptr() for o in outs])

    for _ in range(2):
        call_e2e(ctx, h_out)
    ms_single, wall_single = timed_host_calls(D, lambda: call_e2e(ctx, h_out), args.steps)
    assert np.array_equal(h_out[0].numpy(), res_dev["n_inliers"]), "e2e and resident arms disagree"
    # two host threads, one context each (the library's threading model), alternating steps: one call's H2D overlaps
    # the other's compute tail and D2H.  Every step still moves its inputs and results across PCIe inside the region.
    ctx2 = make_ctx(args, local_rank)
    ctx2.set_option(api.OPT_TC_OPERAND_CACHE, 0)
    h_out2 = [_pinned_zeros(o.numel(), o.numpy().dtype) for o in h_out]

    def streaming(steps):
        n1 = (steps + 1) // 2
        th = [threading.Thread(target=lambda: [call_e2e(ctx, h_out) for _ in range(n1)]),
              threading.Thread(target=lambda: [call_e2e(ctx2, h_out2) for _ in range(steps - n1)])]
        for t in th:
            t.start()
        for t in th:
            t.join()

    streaming(4)
    l1 = ctx.kernel_launches + ctx2.kernel_launches
    ms_stream, wall_stream = timed_host_calls(D, lambda: streaming(args.steps), 1)
    e2e_launches = ctx.kernel_launches + ctx2.kernel_launches - l1
    assert np.array_equal(h_out[0].numpy(), res_dev["n_inliers"]) and (args.steps < 2 or np.array_equal(h_out2[0].numpy(), res_dev["n_inliers"]))
    for k, name in ((3, "train_idx"), (4, "dist"), (5, "mask")):
        assert np.array_equal(h_out[k].numpy(), res_dev[name]), "e2e and resident arms disagree on " + name
    ctx2.close()

    out = dict(kind=kind, P=P, n_kp=n_kp, rot=rot, sc=sc, value=value, ms_per_step=ms_total / args.steps, launches=launches,
               clocks=clocks, ham_ms=float(np.mean(ham_ms)), gms_ms=float(np.mean(gms_ms)), kt=kt, set=s, res=res_dev,
               inliers=D.sum_i(res_dev["n_inliers"].sum()),
               e2e=dict(value=total_pairs * args.steps / (ms_stream * 1e-3), unit="pairs/s", h2d_bytes_per_step=int(h2d),
                        d2h_bytes_per_step=int(d2h), ms_per_step=ms_stream / args.steps, wall_ms_per_step=wall_stream / args.steps,
                        gpu_launches=int(e2e_launches),
                        mode="sfmgms_match_image_set from 2 host threads x 1 context each, alternating steps (one call's H2D overlaps "
                             "the other's compute tail and D2H)",
                        single_thread=dict(value=total_pairs * args.steps / (ms_single * 1e-3), ms_per_step=ms_single / args.steps,
                                           wall_ms_per_step=wall_single / args.steps)))
    ctx.close()
    return out


def check_batch_parity(cpu_res, r, n_kp):
    """GPU train_idx / dist / mask / n_inliers of the sampled pairs == the CPU reference path's, bit for bit"""
    for p, (idx, dist, g) in enumerate(cpu_res):
        sl = slice(p * n_kp, (p + 1) * n_kp)
        if not (np.array_equal(r["train_idx"][sl], idx) and np.array_equal(r["dist"][sl], dist)):
            raise AssertionError("bench parity: BF-Hamming of pair %d differs from the CPU reference path" % p)
        m = r["mask"][sl].astype(bool)
        full = g["mask"] if len(g["mask"]) == n_kp else np.zeros(n_kp, bool)
        if not (np.array_equal(m, full) and r["n_inliers"][p] == g["n_inliers"] and r["mask_len"][p] == len(g["mask"])):
            raise AssertionError("bench parity: GMS mask of pair %d differs from the CPU reference path" % p)
    return len(cpu_res)


def run_allpairs(args, D, local_rank, steps, warmup, n_images=SEQ_IMAGES):
    """config 5, strong scaling: the whole pair list is fixed, rank r owns a contiguous block of the i-major list"""
    import torch

    from sfm_gms_b200 import api
    from sfm_gms_b200 import dist as sd

    rank, world, dev = D.rank, D.world, D.dev
    s = make_sequence(n_images)          # every rank builds the same set (seeded): rank 0's copy is THE input, the
    pairs_all = s["pairs"]               # others keep theirs only to check their own results afterwards
    n_total_pairs = len(pairs_all)
    my_idx, my_pairs = sd.shard_pairs(pairs_all, rank, world, "contiguous")
    n_my = len(my_pairs)
    h_desc = _pinned(s["desc"])
    h_kp = _pinned(s["kp"])
    d_desc = torch.empty_like(h_desc, device=dev)
    d_kp = torch.empty_like(h_kp, device=dev)
    ctx = make_ctx(args, local_rank)
    ctx.set_option(api.OPT_TIMING, 1)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    cap = n_my * SEQ_KP                  # worst case: every match an inlier
    # results: every pair's matchesGMS vector (16-byte cv::DMatch records), counts, winning hypothesis, offsets
    d_m = torch.empty(cap * 4, dtype=torch.int32, device=dev)
    d_ninl, d_bh = (torch.zeros(n_my, dtype=torch.int32, device=dev) for _ in range(2))
    d_off = torch.zeros(n_my + 1, dtype=torch.int64, device=dev)
    h_ninl, h_bh = (_pinned_zeros(n_my, np.int32) for _ in range(2))
    h_off = _pinned_zeros(n_my + 1, np.int64)

    def upload_and_broadcast():
        if rank == 0:
            d_desc.copy_(h_desc, non_blocking=True)
            d_kp.copy_(h_kp, non_blocking=True)
        if world > 1:
            D.dist.broadcast(d_desc, src=0)
            D.dist.broadcast(d_kp, src=0)
        torch.cuda.current_stream().synchronize()

    # ---- value: the set resident in HBM on every GPU (uploaded + broadcast before the timed region) --------------
    upload_and_broadcast()
    ctx.set_images_raw(s["offsets"], d_desc.data_ptr(), d_kp.data_ptr(), s["sizes"], api.SFMGMS_DEVICE, keepalive=(d_desc, d_kp))
    totals = []

    def step_resident():
        totals.append(ctx.match_pairs_compact_raw(my_pairs, 0, 0, 6.0, api.SFMGMS_DEVICE, cap, d_ninl.data_ptr(), d_bh.data_ptr(),
                                                  d_off.data_ptr(), d_m.data_ptr()))

    for _ in range(warmup):
        step_resident()
    # the host result buffer holds exactly what this shard produces (known from the warm-up; the data is deterministic);
    # the device buffer above is sized for the worst case
    h_m = _pinned_zeros(max(1, totals[-1]) * 4, np.int32)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    l0 = ctx.kernel_launches
    ham_ms, gms_ms = [], []

    def step_resident_timed():
        step_resident()
        a, b, _ = ctx.last_timing()
        ham_ms.append(a)
        gms_ms.append(b)

    ms_total = timed_device(D, stream, step_resident_timed, steps)
    launches = ctx.kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    dev_bytes = ctx.device_bytes
    value = n_total_pairs * steps / (ms_total * 1e-3)
    ninl_dev = d_ninl.cpu().numpy().copy()
    host_cap = totals[-1]

    # ---- e2e: host image set -> upload on rank 0 -> NCCL broadcast -> match shard -> results in pinned host memory ----
    bc_ms = []

    def step_e2e():
        t0 = time.perf_counter()
        upload_and_broadcast()
        bc_ms.append(1e3 * (time.perf_counter() - t0))
        ctx.set_images_raw(s["offsets"], d_desc.data_ptr(), d_kp.data_ptr(), s["sizes"], api.SFMGMS_DEVICE, keepalive=(d_desc, d_kp))
        totals.append(ctx.match_pairs_compact_raw(my_pairs, 0, 0, 6.0, api.SFMGMS_HOST, host_cap, h_ninl.data_ptr(), h_bh.data_ptr(),
                                                  h_off.data_ptr(), h_m.data_ptr()))

    for _ in range(max(1, min(warmup, 2))):     # also warms NCCL (communicator set-up happens on the first collective)
        step_e2e()
    del bc_ms[:]
    ms_e2e, wall_e2e = timed_host_calls(D, step_e2e, steps)
    assert np.array_equal(h_ninl.numpy(), ninl_dev), "e2e and resident arms disagree"
    my_inl = int(h_off[n_my].item())
    d2h = my_inl * 16 + n_my * 8 + (n_my + 1) * 8
    h2d = (h_desc.numel() + h_kp.numel() * 4) if rank == 0 else 0

    # ---- secondary e2e: the same step with 8-byte {queryIdx, trainIdx} records (SFMGMS_OPT_COMPACT_RECORD = 1) -- what the
    # reference's consumer of matchesGMS reads; at 8 GPUs the 16-byte records are bound by the host's ingest rate -----------
    hm_dmatch = h_m.numpy().view(api.DMATCH_DT).copy()
    ctx.set_option(api.OPT_COMPACT_RECORD, 1)
    step_e2e()
    ms_e2e8, wall_e2e8 = timed_host_calls(D, step_e2e, max(1, min(steps, 3)))
    ctx.set_option(api.OPT_COMPACT_RECORD, 0)
    ip = h_m.numpy()[: 2 * my_inl].reshape(-1, 2)
    assert np.array_equal(ip[:, 0], hm_dmatch["queryIdx"][:my_inl]) and np.array_equal(ip[:, 1], hm_dmatch["trainIdx"][:my_inl]), \
        "index-pair records differ from the DMatch records"
    e2e8 = dict(value=n_total_pairs * max(1, min(steps, 3)) / (ms_e2e8 * 1e-3), unit="pairs/s", ms_per_step=ms_e2e8 / max(1, min(steps, 3)),
                d2h_bytes_per_step=int(D.sum_i(my_inl * 8 + n_my * 8 + (n_my + 1) * 8)),
                records="8-byte {queryIdx, trainIdx} (SFMGMS_OPT_COMPACT_RECORD = 1); checked equal to the DMatch records' index fields")

    # ---- parity: a seeded sample of THIS rank's pairs against the CPU reference path (after the timed regions) -------
    n_chk = args.check_pairs
    chk = np.sort(np.random.default_rng(777 + rank).choice(n_my, min(n_chk, n_my), replace=False))
    cpu = CpuPath(max(1, (os.cpu_count() or 1) // world))
    hm = hm_dmatch
    ho = h_off.numpy()
    ok = 1
    for p, (idx, dist, g) in zip(chk, cpu_pairs(cpu, s, my_pairs[chk], 0, 0)):
        keep = np.nonzero(g["mask"])[0]
        m = hm[ho[p]:ho[p + 1]]
        if not (np.array_equal(m["queryIdx"], keep) and np.array_equal(m["trainIdx"], idx[keep]) and
                np.array_equal(m["distance"], dist[keep].astype(np.float32)) and h_ninl[p].item() == g["n_inliers"]):
            ok = 0
    if not D.min_i(ok):
        raise AssertionError("bench parity: an all-pairs result differs from the CPU reference path")
    out = dict(value=value, ms_per_step=ms_total / steps, launches=launches, clocks=clocks, ham_ms=float(np.mean(ham_ms)),
               gms_ms=float(np.mean(gms_ms)), n_pairs=n_total_pairs, pairs_per_rank=n_my, inliers=D.sum_i(my_inl),
               parity_checked_pairs=D.sum_i(len(chk)), device_bytes=int(dev_bytes), seq=s, cpu=cpu, e2e_index_pairs=e2e8,
               e2e=dict(value=n_total_pairs * steps / (ms_e2e * 1e-3), unit="pairs/s", h2d_bytes_per_step=int(D.sum_i(h2d)),
                        d2h_bytes_per_step=int(D.sum_i(d2h)), ms_per_step=ms_e2e / steps, wall_ms_per_step=wall_e2e / steps,
                        upload_plus_broadcast_ms=float(np.mean(bc_ms)),
                        mode="per step: pinned host set -> H2D on rank 0 -> NCCL broadcast (desc 164 MB + kp 41 MB) -> "
                             "sfmgms_set_images(DEVICE) -> sfmgms_match_pairs_compact(shard) -> matchesGMS records + counts in this "
                             "rank's pinned host memory"))
    ctx.close()
    return out


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="auto", choices=["auto", "cfg2", "cfg3", "cfg4", "allpairs"])
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step of a batch workload (default: 256 / 32 / 4)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-pairs", type=int, default=8, help="pairs per step of the CPU reference arm")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the cpu_baseline sample (default: sized for ~10 s)")
    ap.add_argument("--check-pairs", type=int, default=4, help="all-pairs: pairs per rank checked against the CPU path")
    ap.add_argument("--allpairs-steps", type=int, default=2, help="steps of the secondary all-pairs measurement at N=1")
    ap.add_argument("--no-allpairs", action="store_true", help="N=1: skip the secondary all-pairs measurement")
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "tc", "fp4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    kind = pick_workload(args, world)
    D = Dist(rank, world, local_rank)
    hkind = "fp4" if args.kernel == "auto" else args.kernel
    threads = os.cpu_count() or 1

    if kind == "allpairs":
        a = run_allpairs(args, D, local_rank, args.steps, args.warmup)
        extra = {}
        if world > 1 and args.workload == "auto":
            # secondary: the round-1 streaming figure (independent cfg2 batches through host buffers, weak scaling)
            b = run_batch_workload(args, D, "cfg2", local_rank)
            extra["e2e_streaming_cfg2"] = dict(b["e2e"], value_resident=b["value"], scaling="weak", pairs_per_gpu_per_step=b["P"])
        if rank == 0:
            dists = float(a["n_pairs"]) * SEQ_KP * SEQ_KP / world        # per rank and step
            roof = hamming_roofline(hkind, dists, a["ham_ms"], a["clocks"])
            line = {"metric": METRIC, "value": a["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": a["ms_per_step"], "us_per_pair_per_gpu": 1e3 * a["ms_per_step"] / a["pairs_per_rank"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                    "dtype": "u8/int32 (f32/f64 cell+threshold decisions)", "data": "synthetic",
                    "config": {"workload": WORKLOADS[kind]},
                    "details": {"hamming_kernel": hkind, "pairs": a["n_pairs"], "pairs_per_gpu": a["pairs_per_rank"],
                                "l2": "per step every GPU streams the whole 205 MB set many times (> 126 MB L2); no flush",
                                "parallelism": "pair-sharded x%d (contiguous blocks of the i-major list), one NCCL broadcast of the "
                                               "shared set per step in the e2e arm, no collective on the result path" % world,
                                "library_device_bytes": a["device_bytes"], "inliers_total": a["inliers"]},
                    "e2e": a["e2e"], "e2e_index_pairs": a["e2e_index_pairs"], "gpu_launches": int(a["launches"]), "roofline": roof,
                    "stage_ms_per_step": {"hamming": a["ham_ms"], "gms": a["gms_ms"]}, "clocks": a["clocks"],
                    "parity_checked_pairs": a["parity_checked_pairs"]}
            line.update(extra)
            if world == 1 and not args.no_cpu_baseline:
                cpu = a["cpu"]
                line["cpu_baseline"] = {"value": cpu.n / (cpu.t_bf + cpu.t_gms), "unit": "pairs/s", "cores": cpu.threads, "kind": "port",
                                        "sample": "%d seeded random pairs of the list (the parity sample); %s" % (cpu.n, cpu.describe())}
            print(json.dumps(line), flush=True)
    else:
        b = run_batch_workload(args, D, kind, local_rank)
        ap_res = None
        if world == 1 and kind == "cfg2" and args.workload == "auto" and not args.no_allpairs:
            ap_res = run_allpairs(args, D, local_rank, max(1, min(args.steps, args.allpairs_steps)), 1)
        if rank == 0:
            P, n_kp = b["P"], b["n_kp"]
            dists = float(P) * n_kp * n_kp
            roof = hamming_roofline(hkind, dists, b["ham_ms"], b["clocks"])
            tr = load_json("profiles", "r2_traffic.json") or load_json("profiles", "r1_traffic.json")
            kname = "hamming_%s" % hkind
            if kname in b["kt"] and b["kt"][kname][0] > 0:
                k_ms = b["kt"][kname][0] / 3.0          # the per-kernel event pass runs 3 steps
                roof["kernel_ms"] = k_ms
                roof["frac_kernel"] = roof["achieved"] * b["ham_ms"] / k_ms / roof["peak"]
                roof["note"] = ("frac = OPs / duration of the whole Hamming STAGE in the timed region (operand unpack + the kernel; the "
                                "kernel settles its own ties); frac_kernel = the same OPs / the kernel's own duration (separate event pass)")
            if kind == "cfg2" and tr.get("kernel") == roof["kernel"]:
                roof["traffic"] = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * P / tr["pairs_per_launch"]
                roof["traffic_source"] = "profiles/%s (ncu dram__bytes_read.sum + dram__bytes_write.sum)" % tr.get("file", "r1_traffic.json")
                # fp4: unpacked train rows (128 B) + packed query rows (32 B, expanded in the kernel) + packed train rows (32 B, the
                # tie resolution's candidates) + keys (4 B); int8: both images unpacked (256 B rows) + keys
                packed_q = hkind == "fp4" and os.environ.get("SFMGMS_FP4_PACKED_QUERIES", "1") != "0"
                roof["algorithmic_bytes_per_launch"] = P * n_kp * (128 + 32 + 32 + 4) if packed_q else \
                    P * (2 * n_kp * (128 if hkind == "fp4" else 256) + 4 * n_kp)
            n_scales = 5 if b["sc"] else 1
            line = {"metric": METRIC, "value": b["value"], "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": b["ms_per_step"], "us_per_pair": 1e3 * b["ms_per_step"] / P, "higher_is_better": True,
                    "scaling": "weak", "vs_baseline": None, "dtype": "u8/int32 (f32/f64 cell+threshold decisions)", "data": "synthetic",
                    "config": {"workload": WORKLOADS[kind]},
                    "details": {"pairs_per_gpu_per_step": P, "hamming_kernel": hkind,
                                "l2": "inputs per step (%.0f MB/GPU) %s the 126 MB L2" % (2 * P * n_kp * 40 / 1e6,
                                                                                        "exceed" if 2 * P * n_kp * 40 > 126e6 else "fit in"),
                                "parallelism": "x%d independent batches, no collective" % world, "inliers_total": b["inliers"]},
                    "e2e": b["e2e"], "gpu_launches": int(b["launches"]), "roofline": roof,
                    "roofline_kernels": kernel_rooflines(b["kt"], 3, hkind, P, n_kp, n_scales),
                    "roofline_gms_stage": gms_stage_roofline(b["gms_ms"], P, n_kp, n_scales, 8 if b["rot"] else 1),
                    "stage_ms_per_step": {"hamming": b["ham_ms"], "gms": b["gms_ms"]}, "clocks": b["clocks"]}
            if world == 1 and not args.no_cpu_baseline:
                # the CPU reference path on a bounded sample of the very pairs the GPU just matched -- its results are
                # the parity check of the GPU arm
                cpu = CpuPath(threads)
                n_s = args.cpu_pairs or {"cfg2": 256, "cfg3": 4, "cfg4": 1}[kind]
                n_s = min(n_s, P)
                res = cpu_pairs(cpu, b["set"], b["set"]["pairs"][:n_s], b["rot"], b["sc"])
                line["parity_checked_pairs"] = check_batch_parity(res, b["res"], n_kp)
                line["cpu_baseline"] = {"value": n_s / (cpu.t_bf + cpu.t_gms), "unit": "pairs/s", "cores": threads, "kind": "port",
                                        "sample": "the first %d pairs of the step's batch; %s" % (n_s, cpu.describe())}
            if ap_res is not None:
                line["allpairs"] = {"workload": WORKLOADS["allpairs"], "value": ap_res["value"], "unit": "pairs/s", "scaling": "strong",
                                    "steps": max(1, min(args.steps, args.allpairs_steps)), "ms_per_step": ap_res["ms_per_step"],
                                    "e2e": ap_res["e2e"], "e2e_index_pairs": ap_res["e2e_index_pairs"], "stage_ms_per_step": {"hamming": ap_res["ham_ms"], "gms": ap_res["gms_ms"]},
                                    "parity_checked_pairs": ap_res["parity_checked_pairs"], "library_device_bytes": ap_res["device_bytes"],
                                    "gpu_launches": int(ap_res["launches"]), "inliers_total": ap_res["inliers"]}
            print(json.dumps(line), flush=True)
    if world > 1:
        D.dist.barrier()
        D.dist.destroy_process_group()


if __name__ == "__main__":
    main()
