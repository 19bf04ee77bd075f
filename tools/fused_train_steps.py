"""Run a few FusedTrainStep steps on the C2 workload (for ncu launch lists / captures of the CUDA-graph training step).
usage: fused_train_steps.py [steps] [N_rays]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200.api import _QueryFn
from nerf_pytorch_b200.trainer import FusedTrainStep
from oracle import synth
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nets = []
for seed in (0, 1):
    m = nb.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.nerf_state(seed).items()}); nets.append(m.to(dev))
e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0); q = _QueryFn(e, ed, 65536, 10, 4, 0)
sb = synth.ray_batch("lego", N, seed=0)
kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=q,
          N_samples=64, N_importance=128, perturb=1., white_bkgd=True, raw_noise_std=0.)
tr = FusedTrainStep(sb["H"], sb["W"], sb["K"], N, kw)
rays = torch.from_numpy(sb["rays"]).pin_memory(); tgt = torch.rand(N, 3).pin_memory()
torch.cuda.synchronize()
print("STEPS-BEGIN", flush=True)
for _ in range(steps):
    loss = tr(rays, tgt)
print("loss", loss)
