"""Time the reference's unfused op chain (oracle/torch_ref.py, stock torch ops) on cuda:0 beside the fused path.

BASELINE.json's target is relative to "the reference PyTorch-GPU rays/sec"; /root/reference does not travel to the
GPU box, so the restatement pinned by tests/test_oracle_golden.py stands in for it.  Same workload as bench.py:
lego 400x400 synthetic rays, 4096 rays, 64 + 128 samples, test-mode kwargs, under no_grad, CUDA-event timed.
Writes gpurun_out/torch_gpu_reference.json.
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, torch_ref as T

dev = torch.device("cuda:0")
N, steps, warm = 4096, 20, 5
sb = synth.ray_batch("lego", N, seed=0)
rays = torch.from_numpy(sb["rays"]).to(dev)
sd = [{k: torch.from_numpy(v).to(dev) for k, v in synth.nerf_state(s).items()} for s in (0, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"workload": "lego 400x400 synthetic rays, N_rand=4096, 64+128 samples, forward, no_grad", "steps": steps, "warmup": warm}
ref_out = None
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    ms = []
    with torch.no_grad():
        for i in range(warm + steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = T.render(rays[0], rays[1], sd[0], sd[1], 2.0, 6.0)
            b.record(); torch.cuda.synchronize()
            if i >= warm: ms.append(a.elapsed_time(b))
    if not tf32: ref_out = r["rgb_map"].clone()
    res["tf32" if tf32 else "fp32"] = {"ms_per_step_median": float(np.median(ms)), "rays_per_s": N / (float(np.median(ms)) * 1e-3)}
# the fused path on the same inputs, for the ratio and a parity figure against the fp32 torch run
import __graft_entry__ as ge; ge.build()
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200.api import _QueryFn
nets = []
for s in (0, 1):
    m = nb.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.nerf_state(s).items()}); nets.append(m.to(dev))
e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0); q = _QueryFn(e, ed, 65536, 10, 4, 0)
kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=q,
          N_samples=64, N_importance=128, perturb=0., white_bkgd=True, raw_noise_std=0.)
ms = []
with torch.no_grad():
    for i in range(warm + steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = nb.render(400, 400, sb["K"], rays=rays, **kw); b.record(); torch.cuda.synchronize()
        if i >= warm: ms.append(a.elapsed_time(b))
res["fused"] = {"ms_per_step_median": float(np.median(ms)), "rays_per_s": N / (float(np.median(ms)) * 1e-3),
                "rgb_rel_l2_vs_torch_fp32": float((out[0] - ref_out).norm() / ref_out.norm())}
res["speedup_vs_torch_fp32"] = res["fused"]["rays_per_s"] / res["fp32"]["rays_per_s"]
res["speedup_vs_torch_tf32"] = res["fused"]["rays_per_s"] / res["tf32"]["rays_per_s"]
res["gpu"] = torch.cuda.get_device_name(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/torch_gpu_reference.json", "w"), indent=1)
print(json.dumps(res))
