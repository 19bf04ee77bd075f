"""Debug: clock64 trace of one super-tile of the fused kernel + tcgen05.mma issue-rate probe."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge; ge.build()
import nerf_pytorch_b200 as nb
from nerf_pytorch_b200 import _lib
from nerf_pytorch_b200.api import _QueryFn
from oracle import synth
lib = _lib.load(); dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
for N in (256, 128):
    for sw64 in (1, 0):
        for reps in (64, 512):
            lib.nerf_b200_debug_mma_rate(reps, N, sw64, C.c_void_p(out.data_ptr()), None); torch.cuda.synchronize()
            o = out.cpu().numpy()
            print(f"mma_rate N={N} B={'SW64' if sw64 else 'SW128'} reps={reps}: total {o[0]} cyc = {o[0]/reps:.1f} cyc/MMA (issue loop {o[1]/reps:.1f} cyc/MMA)")
nets = []
for seed in (0, 1):
    m = nb.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.nerf_state(seed).items()}); nets.append(m.to(dev))
e, _ = nb.get_embedder(10, 0); ed, _ = nb.get_embedder(4, 0); q = _QueryFn(e, ed, 65536, 10, 4, 0)
sb = synth.ray_batch("lego", 4096, seed=0); rays = torch.from_numpy(sb["rays"]).to(dev)
kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=q,
          N_samples=64, N_importance=128, perturb=0., white_bkgd=True, raw_noise_std=0.)
with torch.no_grad():
    for _ in range(3): nb.render(400, 400, sb["K"], rays=rays, **kw)
    tr = torch.zeros(4096, dtype=torch.int64, device=dev)
    lib.nerf_b200_debug_set_trace(C.c_void_p(tr.data_ptr()))
    nb.render(400, 400, sb["K"], rays=rays, **kw); torch.cuda.synchronize()
    lib.nerf_b200_debug_set_trace(None)
t = tr.cpu().numpy()
PAIR = os.environ.get("NERF_B200_PAIR", "1")[:1] != "0"
epi = t[2048:2048 + 128].reshape(2, 16, 4)
if PAIR:
    ps = t[:80].reshape(10, 2, 4); t00 = ps[0, 0, 0]
    print("pair issuer, per pass (layer, slot): start | act_wait | first unit issued after | pass issue time | gap to next pass start")
    flat = [(l, X, ps[l, X]) for l in range(10) for X in range(2) if ps[l, X, 0]]
    for i, (l, X, a) in enumerate(flat):
        nxt = flat[i + 1][2][0] if i + 1 < len(flat) else a[3]
        print(f"  L{l} {'AB'[X]}: start {a[0]-t00:7d}  act_wait {a[1]-a[0]:5d}  first_unit {a[2]-a[1]:5d}  issue {a[3]-a[1]:5d}  gap_next {nxt-a[3]:5d}")
    print("super-tile issuer span:", ps[ps > 0].max() - t00)
else:
    iss = t[:400].reshape(10, 10, 4); prod = t[1024:1024 + 400].reshape(-1, 2)
    t00 = iss[0, 0, 0]
    print("issuer: layer chunk | t_start wfull_wait act_wait(chunk0) chunk_total")
    prev = None
    for l in range(10):
        for c in range(10):
            a = iss[l, c]
            if a[0] == 0: continue
            print(f"  L{l} c{c}: start {a[0]-t00:7d}  wfull_wait {a[1]-a[0]:5d}  act_wait {(a[2]-a[1]) if a[2] else 0:5d}  total {a[3]-a[0]:5d}  gap_prev {(a[0]-prev) if prev else 0:5d}")
            prev = a[3]
    print("super-tile issuer span:", iss[iss > 0].max() - t00)
    pw = [(i, p[1] - p[0], p[0] - t00) for i, p in enumerate(prod) if p[0]]
    print("producer: chunk | wempty_wait")
    print("  ", [(i, int(w)) for i, w, _ in pw][:160])
for X in range(2):
    for l in range(10):
        a = epi[X, l]
        if a[0]: print(f"epi X{X} L{l}: wait_start {a[0]-t00:7d} dfull_wait {a[1]-a[0]:6d} epilogue {a[2]-a[1] if a[2] else -1:6d}")

# per-super-tile timeline of CTA 0 (fine launch): issuer enc_full wait, sampler enc_free wait + encode time
ist = t[2200:2296].reshape(48, 2); smp = t[2300:2396].reshape(48, 2); send = t[2400:2448]
n = int((ist[:, 0] > 0).sum())
print("super-tile | issuer start (rel), encfull_wait | period | sampler encfree_wait, encode")
for s in range(n):
    print(f"  st{s:2d}: start {ist[s,0]-ist[0,0]:8d} encfull_wait {ist[s,1]-ist[s,0]:6d} period {(ist[s,0]-ist[s-1,0]) if s else 0:6d} | "
          f"encfree_wait {smp[s,1]-smp[s,0]:6d} encode {send[s]-smp[s,1]:6d}")

if PAIR:
    w = t[2500:2564].reshape(2, 16, 2)
    print("pair kernel, layer 3 epilogue, per warp (CTA 0 / CTA 1; warps 4..19): fence done, arrive done  (relative to pass L3 A start)")
    base = ps[3, 0, 0]
    for b in range(2):
        print(f"  CTA {b}:", [(int(w[b, i, 0] - base), int(w[b, i, 1] - base)) for i in range(16)])
    print("  issuer: L3 A act seen", int(ps[3,0,1]-base), "L3 A issued", int(ps[3,0,3]-base), "| L3 B act seen", int(ps[3,1,1]-base), "issued", int(ps[3,1,3]-base), "| L4 A act seen", int(ps[4,0,1]-base), "| L4 B act seen", int(ps[4,1,1]-base))
    e = epi; print("  epi warp 4 (CTA 0) L3: d_full seen", int(e[0,3,1]-base), " X1:", int(e[1,3,1]-base))
