import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sfm_gms_b200 as sg
from sfm_gms_b200 import api
import bench
P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
desc, kp = bench.gen_pairs_torch(P, 2, dev)
h_desc = desc.cpu().pin_memory(); h_kp = kp.cpu().pin_memory()
ctx = sg.Context(0)
off = np.arange(2 * P + 1, dtype=np.int64) * bench.N_KP
sizes = np.tile(np.array([[640, 480]], np.int32), (2 * P, 1))
pairs = np.ascontiguousarray(np.arange(2 * P, dtype=np.int32).reshape(-1, 2))
tot = P * bench.N_KP
h = [torch.zeros(P, dtype=torch.int32).pin_memory() for _ in range(3)]
h_ti = torch.zeros(tot, dtype=torch.int32).pin_memory(); h_di = torch.zeros(tot, dtype=torch.int32).pin_memory(); h_mk = torch.zeros(tot, dtype=torch.uint8).pin_memory()
for i in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.match_image_set_raw(off, h_desc.data_ptr(), h_kp.data_ptr(), sizes, pairs, 0, 0, 6.0, h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h_ti.data_ptr(), h_di.data_ptr(), h_mk.data_ptr())
    t1 = time.perf_counter()
    print("call %d: %.3f ms  inliers %d" % (i, 1e3 * (t1 - t0), int(h[0].sum())), flush=True)
