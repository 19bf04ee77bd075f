"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU  --  build container only.

    python oracle/gen_golden.py            # needs /root/reference; writes tests/golden/

Each fixture stores the exact inputs (rays, injected RNG draws) and the reference's outputs.
Network parameters are not stored: they are regenerated from `oracle/synth.nerf_state(seed)`
(deterministic numpy RNG) and loaded into the reference's NeRF modules with load_state_dict.
RNG draws follow the reference's own `pytest=True` hooks (np.random.seed(0) + np.random.rand:
run_nerf.py:373-377, :287-291, run_nerf_helpers.py:210-219), recorded here so that the oracle and
the CUDA path can be fed the same numbers.
"""
import argparse
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def make_args(tmp, N_importance=128, N_samples=64, use_viewdirs=True, white_bkgd=True,
              lindisp=False, dataset_type="blender", no_ndc=False, raw_noise_std=0.0, perturb=1.0):
    os.makedirs(os.path.join(tmp, "exp"), exist_ok=True)
    return types.SimpleNamespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=use_viewdirs,
        N_importance=N_importance, N_samples=N_samples, netdepth=8, netwidth=256,
        netdepth_fine=8, netwidth_fine=256, netchunk=1024 * 64, lrate=5e-4,
        basedir=tmp, expname="exp", ft_path=None, no_reload=True, perturb=perturb,
        white_bkgd=white_bkgd, raw_noise_std=raw_noise_std, dataset_type=dataset_type,
        no_ndc=no_ndc, lindisp=lindisp)


def load_state(model, state):
    model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in state.items()})


def to_np(x):
    return x.detach().cpu().numpy()


def run_case(rn, name, scene, N, N_importance, seed_w, sharpen=False, perturb=0.0, pytest=False,
             lindisp=False, raw_noise_std=0.0, grads=False, ray_seed=0):
    sb = synth.ray_batch(scene, N, seed=ray_seed)
    with tempfile.TemporaryDirectory() as tmp:
        args = make_args(tmp, N_importance=N_importance, white_bkgd=sb["white_bkgd"], lindisp=lindisp,
                         dataset_type="llff" if sb["ndc"] else "blender",
                         raw_noise_std=raw_noise_std, perturb=perturb)
        kw_train, kw_test, start, grad_vars, optim = rn.create_nerf(args)
    load_state(kw_train["network_fn"], synth.nerf_state(seed_w, sharpen))
    if kw_train["network_fine"] is not None:
        load_state(kw_train["network_fine"], synth.nerf_state(seed_w + 1, sharpen))
    kw = dict(kw_train)
    kw.update(near=sb["near"], far=sb["far"], perturb=perturb, raw_noise_std=raw_noise_std)
    rays = torch.from_numpy(sb["rays"])
    K = sb["K"]
    fx = dict(rays=sb["rays"], K=K, H=sb["H"], W=sb["W"], near=sb["near"], far=sb["far"],
              ndc=sb["ndc"], white_bkgd=sb["white_bkgd"], N_samples=64, N_importance=N_importance,
              seed_w=seed_w, sharpen=sharpen, perturb=perturb, lindisp=lindisp,
              raw_noise_std=raw_noise_std, scene=scene)
    S_f = 64 + N_importance
    if pytest:
        # what the reference's pytest hooks will draw (same seed before every draw)
        if perturb > 0:
            np.random.seed(0); fx["t_rand"] = np.random.rand(N, 64).astype(np.float32)
            np.random.seed(0); fx["u"] = np.random.rand(N, max(N_importance, 1)).astype(np.float32)[:, :N_importance]
        if raw_noise_std > 0:
            np.random.seed(0); fx["noise0"] = (np.random.rand(N, 64) * raw_noise_std).astype(np.float32)
            np.random.seed(0); fx["noise1"] = (np.random.rand(N, S_f) * raw_noise_std).astype(np.float32)
    ctx = torch.enable_grad() if grads else torch.no_grad()
    with ctx:
        rgb, disp, acc, extras = rn.render(sb["H"], sb["W"], K, chunk=1024 * 32, rays=rays,
                                           retraw=True, pytest=pytest, **kw)
        fx.update(rgb_map=to_np(rgb), disp_map=to_np(disp), acc_map=to_np(acc),
                  raw=to_np(extras["raw"]))
        for k in ("rgb0", "disp0", "acc0", "z_std"):
            if k in extras:
                fx[k] = to_np(extras[k])
        if grads:
            target = torch.from_numpy(np.random.default_rng(5).random((N, 3), dtype=np.float32))
            fx["target"] = to_np(target)
            loss = torch.mean((rgb - target) ** 2)
            if "rgb0" in extras:
                loss = loss + torch.mean((extras["rgb0"] - target) ** 2)     # run_nerf.py:765-772
            loss.backward()
            fx["loss"] = np.float32(loss.item())
            rng = np.random.default_rng(9)
            for tag, net in (("c", kw_train["network_fn"]), ("f", kw_train["network_fine"])):
                if net is None:
                    continue
                for pname, prm in net.named_parameters():
                    g = to_np(prm.grad).reshape(-1)
                    idx = rng.integers(0, g.size, min(256, g.size))
                    fx[f"g_{tag}_{pname}_idx"] = idx.astype(np.int64)
                    fx[f"g_{tag}_{pname}_val"] = g[idx]
                    fx[f"g_{tag}_{pname}_norm"] = np.float32(np.linalg.norm(g.astype(np.float64)))
                    fx[f"g_{tag}_{pname}_sum"] = np.float32(g.astype(np.float64).sum())
    # intermediate tensors, by re-running the reference's own pieces (for kernel-level parity)
    with torch.no_grad():
        all_ret = None
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **fx)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in fx.items() if not k.startswith("g_")})


def run_units(rn, rh):
    """Function-level vectors: embed, NeRF.forward, raw2outputs, sample_pdf."""
    rng = np.random.default_rng(3)
    fx = {}
    x = (rng.random((64, 3), dtype=np.float32) * 8 - 4)
    fx["embed_x"] = x
    for L in (10, 4, 2):
        fn, od = rh.get_embedder(L, 0)
        fx[f"embed_L{L}"] = to_np(fn(torch.from_numpy(x)))
    # NeRF.forward
    model = rh.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
    load_state(model, synth.nerf_state(7))
    xin = np.concatenate([to_np(rh.get_embedder(10, 0)[0](torch.from_numpy(x * 0.9))),
                          to_np(rh.get_embedder(4, 0)[0](torch.from_numpy(x / np.linalg.norm(x, axis=-1, keepdims=True))))], -1)
    fx["nerf_x"] = xin.astype(np.float32)
    with torch.no_grad():
        fx["nerf_y"] = to_np(model(torch.from_numpy(fx["nerf_x"])))
    # raw2outputs
    raw = (rng.standard_normal((48, 40, 4)).astype(np.float32) * 2)
    raw[5, :, 3] = -1.0                      # all sigma <= 0  -> disp NaN
    raw[6, -1, 3] = 1e-9                     # last-interval saturation
    z = np.sort(rng.random((48, 40), dtype=np.float32) * 4 + 2, -1)
    d = rng.standard_normal((48, 3)).astype(np.float32)
    fx.update(r2o_raw=raw, r2o_z=z, r2o_d=d)
    for wb in (False, True):
        with torch.no_grad():
            outs = rn.raw2outputs(torch.from_numpy(raw), torch.from_numpy(z), torch.from_numpy(d), 0, wb)
        for nm, o in zip(("rgb", "disp", "acc", "weights", "depth"), outs):
            fx[f"r2o_{nm}_wb{int(wb)}"] = to_np(o)
    np.random.seed(0); noise = (np.random.rand(48, 40) * 0.7).astype(np.float32)
    fx["r2o_noise"] = noise
    with torch.no_grad():
        outs = rn.raw2outputs(torch.from_numpy(raw), torch.from_numpy(z), torch.from_numpy(d), 0.7, True, pytest=True)
    for nm, o in zip(("rgb", "disp", "acc", "weights", "depth"), outs):
        fx[f"r2o_{nm}_noise"] = to_np(o)
    # sample_pdf
    bins = np.sort(rng.random((40, 63), dtype=np.float32) * 4 + 2, -1)
    w = rng.random((40, 62), dtype=np.float32) ** 4
    w[3] = 0.0                               # zero weights -> uniform
    w[4, :] = 0.0; w[4, 17] = 1.0            # a spike
    fx.update(spdf_bins=bins, spdf_w=w)
    with torch.no_grad():
        fx["spdf_det"] = to_np(rh.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), 128, det=True))
        fx["spdf_rand"] = to_np(rh.sample_pdf(torch.from_numpy(bins), torch.from_numpy(w), 128, det=False, pytest=True))
    np.random.seed(0); fx["spdf_u"] = np.random.rand(40, 128).astype(np.float32)
    # rays
    H, W, K, c2w = synth.lego_camera(40)
    o, dd = rh.get_rays(H, W, K, torch.from_numpy(c2w))
    fx.update(rays_K=K, rays_c2w=c2w, rays_o=to_np(o), rays_d=to_np(dd))
    Hf, Wf, Kf, c2wf = 38, 50, synth.intrinsics(38, 50, 40.7), synth.fern_camera()[3]
    o, dd = rh.get_rays(Hf, Wf, Kf, torch.from_numpy(c2wf))
    no, nd = rh.ndc_rays(Hf, Wf, Kf[0][0], 1.0, o, dd)
    fx.update(ndc_K=Kf, ndc_o=to_np(no), ndc_d=to_np(nd))
    np.savez_compressed(os.path.join(OUT, "units.npz"), **fx)
    print("wrote units")


def main():
    ap = argparse.ArgumentParser()
    ap.parse_args()
    assert ref_import.available(), "reference not found (this script only runs in the build container)"
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    rn, rh = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    run_units(rn, rh)
    run_case(rn, "lego_det", "lego", 48, 128, seed_w=0)
    run_case(rn, "lego_sharp_det", "lego", 48, 128, seed_w=2, sharpen=True)
    run_case(rn, "lego_perturb", "lego", 48, 128, seed_w=0, perturb=1.0, pytest=True)
    run_case(rn, "lego_coarse_only", "lego", 64, 0, seed_w=4)
    run_case(rn, "lego_lindisp", "lego", 32, 128, seed_w=0, lindisp=True)
    run_case(rn, "fern_ndc_det", "fern", 48, 128, seed_w=6)
    run_case(rn, "fern_ndc_noise", "fern", 32, 64, seed_w=6, perturb=1.0, pytest=True, raw_noise_std=1.0)
    run_case(rn, "lego_grads", "lego", 24, 128, seed_w=0, grads=True)
    run_case(rn, "lego_sharp_grads", "lego", 24, 128, seed_w=2, sharpen=True, grads=True)


if __name__ == "__main__":
    main()
