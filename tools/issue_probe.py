import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
from nerf_pytorch_b200 import _lib
lib = _lib.load_dev(); dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
reps = 256
for nmma in (1, 2, 4, 16):
    for flags, name in ((0, "divergent: mma only"), (3, "divergent: +try_wait+commit"), (128, "uniform: mma only"), (128 + 3, "uniform: +try_wait+commit"), (128 + 11, "uniform: +try_wait+fence+commit")):
        lib.nerf_b200_debug_issue_probe(reps, nmma, flags, C.c_void_p(out.data_ptr()), None); torch.cuda.synchronize()
        o = out.cpu().numpy()
        print(f"nmma={nmma:2d} {name:34s}: {o[0]/reps/nmma:7.1f} cyc/MMA ({o[0]/reps:7.1f} per iter)")
