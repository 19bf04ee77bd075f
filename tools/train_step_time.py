"""Time one training step (forward render + loss + backward + Adam) on the C2 workload."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200.api import _QueryFn
from oracle import synth
dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
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
for _ in range(2): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
K = 3
for _ in range(K): l = step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
print(f"train step N={N}: {dt*1e3:.1f} ms  -> {N/dt:.0f} rays/s  (loss {float(l):.4f})")
