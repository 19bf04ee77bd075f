"""Trained-like golden fixtures  --  build container only (needs /root/reference).

    python oracle/gen_trained.py [steps]          # train, then write every fixture
    python oracle/gen_trained.py fixtures         # re-generate the fixtures from the stored tests/golden/trained_weights.npz

1. Trains the UNMODIFIED reference (run_nerf.create_nerf / render / img2mse, torch.optim.Adam, the reference's lr decay) on
   the CPU for a few hundred steps on an analytic scene -- a shaded unit sphere in front of a white background, seen from
   cameras on the lego orbit (load_blender.pose_spherical) -- so that the networks have STRUCTURE: acc_map spans 0..1,
   weights are peaked, rgb varies.  The default-initialised fixtures of gen_golden.py render an almost constant image,
   which makes a relative-L2 gate on rgb_map easy to pass (round-1 verdict).
2. Stores the trained weights (tests/golden/trained_weights.npz) and the reference's outputs on them:
      trained_lego_1024 : 1024 lego rays, deterministic; rgb/disp/acc/rgb0/disp0/acc0/z_std, raw, and -- captured from the
                          reference's own raw2outputs / sample_pdf calls -- coarse and fine `weights` and `z_vals`
      trained_lego_grads: 256 rays, all 48 gradient tensors of the reference's loss (sampled entries + norms)
      fern_ndc_4096     : BASELINE config 3's shape (4096 fern NDC rays, 64 + 128 samples) on the SAME trained networks (the
                          default initialisation renders acc0 == 0 on NDC rays, which makes coarse parity vacuous), outputs only
   Rays are regenerated from oracle/synth (seeded), so only outputs are stored.
"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, synth  # noqa: E402
from oracle.gen_golden import make_args, load_state, to_np  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sphere_targets(rays_o, rays_d):
    """analytic render of a unit sphere at the origin, white background: [N,3] in [0,1]"""
    d = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    b = np.sum(rays_o * d, -1)
    c = np.sum(rays_o * rays_o, -1) - 1.0
    disc = b * b - c
    hit = disc > 0
    t = -b - np.sqrt(np.where(hit, disc, 0.0))
    p = rays_o + t[:, None] * d
    n = p                                             # unit sphere: normal = position
    col = 0.5 + 0.5 * n
    col = col * (0.6 + 0.4 * np.sign(np.sin(6 * p[:, 0:1]) * np.sin(6 * p[:, 1:2]) * np.sin(6 * p[:, 2:3])))   # checker
    return np.where(hit[:, None], np.clip(col, 0, 1), 1.0).astype(np.float32)


def train(rn, steps, n_rand=512, lrate=1e-3, seed=0):
    H, W, K, _ = synth.lego_camera(100)
    with tempfile.TemporaryDirectory() as tmp:
        args = make_args(tmp, perturb=1.0)
        args.lrate = lrate
        tr, te, start, grad_vars, opt = rn.create_nerf(args)
    load_state(tr["network_fn"], synth.nerf_state(20))
    load_state(tr["network_fine"], synth.nerf_state(21))
    tr.update(near=2., far=6.)
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    t0 = time.time()
    for i in range(steps):
        c2w = synth.pose_spherical(rng.uniform(-180, 180), rng.uniform(-70, -10), 4.0)[:3, :4]
        o, d = synth.camera_rays(H, W, K, c2w)
        idx = rng.permutation(H * W)[:n_rand]
        ro, rd = o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]
        target = torch.from_numpy(sphere_targets(ro, rd))
        rays = torch.from_numpy(np.stack([ro, rd], 0))
        rgb, disp, acc, extras = rn.render(H, W, K, chunk=32768, rays=rays, verbose=False, retraw=True, **tr)   # run_nerf.py:760
        opt.zero_grad()
        loss = rn.img2mse(rgb, target) + rn.img2mse(extras["rgb0"], target)                                      # :764-772
        loss.backward()
        opt.step()
        new_lrate = lrate * (0.1 ** (i / (250 * 1000)))                                                          # :779-783
        for pg in opt.param_groups:
            pg["lr"] = new_lrate
        if i % 20 == 0 or i == steps - 1:
            print(f"step {i}: loss {loss.item():.5f} acc [{acc.min().item():.3f}, {acc.max().item():.3f}] mean {acc.mean().item():.3f}  ({time.time() - t0:.0f} s)", flush=True)
    return tr, te


def spy_render(rn, rh, H, W, K, rays, kw, grads=False):
    """reference render() with its raw2outputs calls recorded (weights / z_vals are not returned by render_rays)"""
    rec = []
    orig = rn.raw2outputs

    def spy(raw, z_vals, rays_d, *a, **k):
        out = orig(raw, z_vals, rays_d, *a, **k)
        rec.append((to_np(z_vals), to_np(out[3]), to_np(raw)))
        return out
    rn.raw2outputs = spy
    try:
        ctx = torch.enable_grad() if grads else torch.no_grad()
        with ctx:
            out = rn.render(H, W, K, chunk=32768, rays=rays, retraw=True, **kw)
    finally:
        rn.raw2outputs = orig
    return out, rec


def main():
    assert ref_import.available(), "reference not found (this script only runs in the build container)"
    torch.set_num_threads(os.cpu_count())
    rn, rh = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "fixtures":
        with tempfile.TemporaryDirectory() as tmp:
            tr, te, _, _, _ = rn.create_nerf(make_args(tmp, perturb=1.0))
        tr.update(near=2., far=6.)
        stored = np.load(os.path.join(OUT, "trained_weights.npz"))
        for tag, key in (("c", "network_fn"), ("f", "network_fine")):
            load_state(tr[key], {k[2:]: stored[k] for k in stored.files if k.startswith(tag + ".")})
    else:
        steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
        tr, te = train(rn, steps)
        w = {}
        for tag, key in (("c", "network_fn"), ("f", "network_fine")):
            for k, v in tr[key].state_dict().items():
                w[f"{tag}.{k}"] = to_np(v)
        np.savez_compressed(os.path.join(OUT, "trained_weights.npz"), steps=steps, **w)
    kw = dict(te); kw.update(near=2., far=6.)
    # ---- 1024 rays, deterministic, with weights / z captured ----
    sb = synth.ray_batch("lego", 1024, seed=11)
    (rgb, disp, acc, ex), rec = spy_render(rn, rh, sb["H"], sb["W"], sb["K"], torch.from_numpy(sb["rays"]), kw)
    fx = dict(N=1024, ray_seed=11, rgb_map=to_np(rgb), disp_map=to_np(disp), acc_map=to_np(acc), rgb0=to_np(ex["rgb0"]), disp0=to_np(ex["disp0"]),
              acc0=to_np(ex["acc0"]), z_std=to_np(ex["z_std"]), z_coarse=rec[0][0], w_coarse=rec[0][1], raw_coarse=rec[0][2][:256],
              z_fine=rec[1][0], w_fine=rec[1][1], raw_fine=rec[1][2][:256])          # raw: first 256 rays (fixture size)
    np.savez_compressed(os.path.join(OUT, "trained_lego_1024.npz"), **fx)
    print("trained_lego_1024: acc range", fx["acc_map"].min(), fx["acc_map"].max(), "rgb std", fx["rgb_map"].std())
    # ---- gradients on the trained weights ----
    sb = synth.ray_batch("lego", 256, seed=12)
    kwt = dict(tr); kwt.update(perturb=0.)
    for key in ("network_fn", "network_fine"):
        for p in kwt[key].parameters():
            p.grad = None
    with torch.enable_grad():
        rgb, disp, acc, ex = rn.render(sb["H"], sb["W"], sb["K"], chunk=32768, rays=torch.from_numpy(sb["rays"]), retraw=True, **kwt)
        target = torch.from_numpy(sphere_targets(sb["rays"][0], sb["rays"][1]))
        loss = rn.img2mse(rgb, target) + rn.img2mse(ex["rgb0"], target)
        loss.backward()
    fx = dict(N=256, ray_seed=12, target=to_np(target), loss=np.float32(loss.item()), rgb_map=to_np(rgb), rgb0=to_np(ex["rgb0"]))
    rng = np.random.default_rng(9)
    for tag, key in (("c", "network_fn"), ("f", "network_fine")):
        for pname, prm in kwt[key].named_parameters():
            g = to_np(prm.grad).reshape(-1)
            idx = rng.integers(0, g.size, min(512, g.size))
            fx[f"g_{tag}_{pname}_idx"] = idx.astype(np.int64)
            fx[f"g_{tag}_{pname}_val"] = g[idx]
            fx[f"g_{tag}_{pname}_norm"] = np.float32(np.linalg.norm(g.astype(np.float64)))
    np.savez_compressed(os.path.join(OUT, "trained_lego_grads.npz"), **fx)
    # ---- BASELINE config 3 shape: 4096 fern NDC rays on the trained networks ----
    sb = synth.ray_batch("fern", 4096, seed=13)
    kte = dict(te)
    kte.update(near=0., far=1., ndc=True, white_bkgd=False)
    kte.pop("lindisp", None)
    (rgb, disp, acc, ex), rec = spy_render(rn, rh, sb["H"], sb["W"], sb["K"], torch.from_numpy(sb["rays"]), kte)
    fx = dict(N=4096, ray_seed=13, weights="trained", rgb_map=to_np(rgb), disp_map=to_np(disp), acc_map=to_np(acc), rgb0=to_np(ex["rgb0"]), disp0=to_np(ex["disp0"]),
              acc0=to_np(ex["acc0"]), z_std=to_np(ex["z_std"]))
    np.savez_compressed(os.path.join(OUT, "fern_ndc_4096.npz"), **fx)
    print("fern_ndc_4096: acc0 mean", fx["acc0"].mean(), "std", fx["acc0"].std(), "acc mean", fx["acc_map"].mean())
    print("done")


if __name__ == "__main__":
    main()
