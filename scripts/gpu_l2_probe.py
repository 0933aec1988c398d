"""TMA fill bandwidth from L2: unicast vs cluster multicast (see csrc/probe.cu)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from jimm_b200 import _lib
lib = _lib.load()
rows = 64 * 1024 * 1024 // 128   # 64 MB buffer: L2 resident
buf = torch.randn(rows, 64, device="cuda").half()
iters = 4000
sms = torch.cuda.get_device_properties(0).multi_processor_count
for cluster, mode in ((1, 0), (2, 0), (2, 1), (2, 2), (4, 1), (4, 2), (8, 1), (8, 2)):
    ms = C.c_float()
    rc = lib.jimm_k_l2_probe(C.c_void_p(buf.data_ptr()), rows, mode, cluster, iters, C.byref(ms), None)
    if rc:
        print(cluster, mode, "ERR", lib.jimm_last_error().decode()); continue
    grid = sms // cluster * cluster
    recv = grid * iters * 16384 / (ms.value * 1e-3) / 1e12
    l2 = recv if mode != 2 else recv / cluster
    print(f"cluster {cluster} mode {mode} ({['own tiles','same tile, all load','same tile, multicast'][mode]:22s}): {ms.value:8.3f} ms  received {recv:6.2f} TB/s  ({recv*1e12/grid/1.9e9:5.1f} B/clk/SM @1.9GHz)  L2 reads {l2:6.2f} TB/s", flush=True)
