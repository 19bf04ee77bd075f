import ctypes as C, os, sys, torch
sys.path.insert(0, os.getcwd())
import __graft_entry__ as ge; ge.build()
from nerf_pytorch_b200 import _lib
lib = _lib.load_dev(); dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
reps = 256
for nmma in (2, 4, 8):
    for flags, name in ((128, "mma only"), (128+1, "+wait"), (128+2, "+commit"), (128+6, "+2 commits"), (128+3, "+wait+commit"), (128+11, "+wait+fence+commit"), (128+9, "+wait+fence")):
        lib.nerf_b200_debug_issue_probe(reps, nmma, flags, C.c_void_p(out.data_ptr()), None); torch.cuda.synchronize()
        o = out.cpu().numpy()
        print(f"nmma={nmma:2d} {name:22s}: {o[0]/reps:7.1f} per iter (issue loop {o[1]/reps:7.1f}) ideal {nmma*128.8:6.1f}")
