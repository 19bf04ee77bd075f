import cProfile, pstats, sys, os, io
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, numpy as np
import gpu_common as G
sb = G.synth.ray_batch("lego", 4096, seed=0)
nets = [G.make_net(G.synth.nerf_state(0)), G.make_net(G.synth.nerf_state(1))]
kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=G.query_fn(),
          N_samples=64, N_importance=128, perturb=0., white_bkgd=True, raw_noise_std=0.)
rays = G.dev(sb["rays"])
def run(n):
    with torch.no_grad():
        for _ in range(n):
            r = G.nb.render(400, 400, sb["K"], rays=rays, **kw)
    torch.cuda.synchronize()
run(20)
import time
t=time.perf_counter(); run(200); print("ms per call", (time.perf_counter()-t)*1000/200)
# host-only time: no sync per call, measure enqueue rate
pr = cProfile.Profile(); pr.enable(); run(300); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
