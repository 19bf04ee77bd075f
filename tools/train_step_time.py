"""Time one training step (run_nerf.py:760-776: render with retraw, two MSE terms, backward, Adam) on the C2 workload through
the public API (eager autograd path), CUDA-event timed, with the device time of the three tensor-core kernels split out.
Writes gpurun_out/train_step.json.   usage: train_step_time.py [N_rays] [steps] [backward: tc|exact]"""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200 import _lib
from nerf_pytorch_b200.api import _QueryFn
from oracle import synth
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nb.set_backward(sys.argv[3] if len(sys.argv) > 3 else "tc")
lib = _lib.load()
nets = []
for seed in (0, 1):
    m = nb.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.nerf_state(seed).items()}); nets.append(m.to(dev))
e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0); q = _QueryFn(e, ed, 65536, 10, 4, 0)
sb = synth.ray_batch("lego", N, seed=0); rays = torch.from_numpy(sb["rays"]).to(dev)
target = torch.rand(N, 3, device=dev)
opt = torch.optim.Adam(list(nets[0].parameters()) + list(nets[1].parameters()), lr=5e-4)
kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=q,
          N_samples=64, N_importance=128, perturb=1., white_bkgd=True, raw_noise_std=0., retraw=True)
def step():
    rgb, disp, acc, ex = nb.render(400, 400, sb["K"], rays=rays, **kw)
    opt.zero_grad()
    loss = nb.img2mse(rgb, target) + nb.img2mse(ex["rgb0"], target)
    loss.backward(); opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
l0 = nb.launch_count(); step(); launches = nb.launch_count() - l0
lib.nerf_b200_timing_enable(1)
ms = []
for _ in range(K):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); l = step(); b.record(); torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
kms, kn, kfl = C.c_double(), C.c_int64(), C.c_double()
kinds = (C.c_double * 3)()
lib.nerf_b200_timing_read_kinds(C.byref(kms), C.byref(kn), C.byref(kfl), kinds)
lib.nerf_b200_timing_enable(0)
med = float(np.median(ms))
res = {"N": N, "steps": K, "backward": nb.get_backward(), "ms_per_step_median": med, "ms_per_step_min": float(min(ms)), "rays_per_s": N / (med * 1e-3),
       "library_launches_per_step": int(launches), "loss": float(l),
       "kernel_ms_per_step": {"forward_passes": kinds[0] / K, "dgrad_chains": kinds[1] / K, "wgrad": kinds[2] / K},
       "train_flop_per_ray": 893190144, "tflops_algorithmic": N * 893190144 / (med * 1e-3) / 1e12,
       "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/train_step.json", "w"), indent=1)
print(json.dumps(res))
