import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
from nerf_pytorch_b200 import _lib
lib = _lib.load_dev(); dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
names = {0: "32x32b.x32 (1 in flight)", 1: "32x32b.x32 x2 per wait", 3: "16x256b.x8 x2 per wait", 4: "16x128b.x16 x2 per wait", 5: "32x32b.x16 x4 per wait"}
reps = 64
for mma in (0, 1):
    for nw in (8, 4):
        for shape in (0, 1, 3, 4, 5):
            lib.nerf_b200_debug_ldtm_rate(reps, shape, nw, mma, C.c_void_p(out.data_ptr()), None); torch.cuda.synchronize()
            cyc = out[0].item() / reps
            print(f"mma={mma} warps={nw} {names[shape]:28s}: {cyc:7.0f} cyc per 128x256 fp32 drain = {131072/cyc:6.1f} B/cyc")
