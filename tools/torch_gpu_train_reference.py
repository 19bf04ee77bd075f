"""Time the reference's TRAIN step (render with retraw, two MSE terms, backward, Adam: run_nerf.py:760-776) as stock torch
ops on cuda:0 (oracle/torch_ref.py restatement; /root/reference does not travel to the GPU box), fp32 and TF32 matmuls.
Writes gpurun_out/torch_gpu_train_reference.json."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, torch_ref as T

dev = torch.device("cuda:0")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps, warm = 8, 3
sb = synth.ray_batch("lego", N, seed=0)
rays = torch.from_numpy(sb["rays"]).to(dev)
target = torch.rand(N, 3, device=dev)
res = {"workload": f"lego 400x400 synthetic rays, N_rand={N}, 64+128 samples, train step (fwd + 2 MSE + bwd + Adam)", "steps": steps, "warmup": warm}
for tf32 in (False, True):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    sd = [{k: torch.from_numpy(v).to(dev).requires_grad_(True) for k, v in synth.nerf_state(s).items()} for s in (0, 1)]
    opt = torch.optim.Adam([p for d in sd for p in d.values()], lr=5e-4, betas=(0.9, 0.999))
    ms = []
    for i in range(warm + steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = T.render(rays[0], rays[1], sd[0], sd[1], 2.0, 6.0)
        opt.zero_grad()
        loss = torch.mean((r["rgb_map"] - target) ** 2) + torch.mean((r["rgb0"] - target) ** 2)
        loss.backward()
        opt.step()
        b.record(); torch.cuda.synchronize()
        if i >= warm: ms.append(a.elapsed_time(b))
    res["tf32" if tf32 else "fp32"] = {"ms_per_step_median": float(np.median(ms)), "rays_per_s": N / (float(np.median(ms)) * 1e-3), "loss": float(loss)}
    del sd, opt, r, loss
    torch.cuda.empty_cache()
res["gpu"] = torch.cuda.get_device_name(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/torch_gpu_train_reference.json", "w"), indent=1)
print(json.dumps(res))
