import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
from nerf_pytorch_b200 import _lib
lib = _lib.load_dev(); dev = torch.device("cuda:0")
buf = torch.zeros(1179648, dtype=torch.uint8, device=dev)       # the packed weight stream size (1152 KB)
out = torch.zeros(256, dtype=torch.int64, device=dev)
for nb in (148, 74, 16, 1):
    for chunk, stages in ((16384, 3), (8192, 6), (16384, 6), (8192, 12), (32768, 3)):
        passes = 8
        lib.nerf_b200_debug_l2_stream(C.c_void_p(buf.data_ptr()), buf.numel(), chunk, stages, passes, nb, C.c_void_p(out.data_ptr()), None)
        torch.cuda.synchronize()
        cyc = out[:nb].float().mean().item()
        tot = (buf.numel() // chunk) * chunk * passes
        print(f"blocks={nb:3d} chunk={chunk:5d} stages={stages:2d}: {tot/cyc:6.1f} B/cycle/SM  ({tot/cyc*nb/1024:6.2f} KB/cycle chip)  {cyc/(tot/chunk):7.0f} cyc/chunk")
