"""Sweep NERF_B200_BWD_OVERLAP ("w[,chunks]": SMs given to the weight gradient that runs next to the data-gradient chain, and the
number of chunks the fine pass is cut into; 0 = passes one after the other) on the C2 training step (FusedTrainStep, one CUDA graph
per step, device-resident batch, CUDA events, L2 not flushed).  One sub-process per setting (the library reads the variable once).
usage: bwd_overlap_sweep.py [setting ...]      e.g.  0  48,1  56,2  64,3"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json, statistics
import numpy as np, torch
sys.path.insert(0, %r)
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200.api import _QueryFn
from nerf_pytorch_b200.trainer import FusedTrainStep
from oracle import synth
dev = torch.device("cuda:0"); N = 4096
nets = []
for seed in (0, 1):
    m = nb.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.nerf_state(seed).items()}); nets.append(m.to(dev))
e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0); q = _QueryFn(e, ed, 65536, 10, 4, 0)
sb = synth.ray_batch("lego", N, seed=0)
kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=q,
          N_samples=64, N_importance=128, perturb=1., white_bkgd=True, raw_noise_std=0.)
tr = FusedTrainStep(sb["H"], sb["W"], sb["K"], N, kw)
tr.rays.copy_(torch.from_numpy(sb["rays"])); tr.target.copy_(torch.rand(N, 3))
for _ in range(5): tr.step_device()
torch.cuda.synchronize()
ms = []
for _ in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tr.step_device(); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
g = tr.flat_g.double()
print(json.dumps({"setting": os.environ.get("NERF_B200_BWD_OVERLAP", "(default)"), "ms_median": statistics.median(ms), "ms_min": min(ms),
                  "loss": float(tr.state[0]), "grad_l2": float(g.norm()), "grad_finite": bool(torch.isfinite(g).all())}))
''' % ROOT

settings = sys.argv[1:] or ["0", "40,1", "48,1", "56,1", "64,1", "48,2", "56,2", "64,2", "72,2", "56,3", "64,3", "64,4"]
for s in settings:
    env = dict(os.environ)
    if s == "default": env.pop("NERF_B200_BWD_OVERLAP", None)
    else: env["NERF_B200_BWD_OVERLAP"] = s
    r = subprocess.run(["timeout", "120", sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    print(line[-1] if line else json.dumps({"setting": s, "error": (r.stderr or r.stdout)[-400:]}), flush=True)
