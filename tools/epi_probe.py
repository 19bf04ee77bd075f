import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
from nerf_pytorch_b200 import _lib
lib = _lib.load_dev(); dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
reps = 64
for mma in (0, 1):
    for mode, name in ((0, "ld only"), (4, "ld, no proxy fence"), (12, "ld, no fences"), (28, "ld, no fences/bar"), (32 + 28, "dual ld, no fences/bar"), (32, "dual ld"),
                       (2, "full"), (2 + 4, "full, no proxy fence"), (2 + 28, "full, no fences/bar")):
        lib.nerf_b200_debug_epi_rate(reps, mode, mma, C.c_void_p(out.data_ptr()), None); torch.cuda.synchronize()
        o = out.cpu().numpy()
        print(f"mma={mma} {name:26s}: {o[0]/reps:7.0f} cyc per tile-layer; ld+wait total {o[1]/reps:7.0f} cyc per rep")
