"""EXPERIMENTAL: building blocks of the tensor-core backward (nerf-pytorch_b200/csrc/bwd_tc.cuh).

Written at the end of round 1 after the GPU budget was spent, so these kernels have never run on a GPU.  The tests are
skipped unless NERF_B200_EXPERIMENTAL=1; round 2 starts by making them pass.  Nothing on a default path uses them."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("NERF_B200_EXPERIMENTAL") != "1", reason="unvalidated round-2 groundwork; set NERF_B200_EXPERIMENTAL=1")]


@pytest.fixture(scope="module")
def G():
    import gpu_common
    return gpu_common


def _img_bytes(M, C):
    return ((M + 127) // 128) * (C // 64) * 16384


def _pack(G, lib, x, scale=1.0):
    M, C = x.shape
    img = torch.zeros(_img_bytes(M, C), dtype=torch.uint8, device=G.DEV)
    xd = G.dev(x)
    G._lib.check(lib.nerf_b200_exp_tile_pack(G.ptr(xd), M, C, scale, G.ptr(img), G.stream()), "tile_pack")
    return img


def _unpack(G, lib, img, M, C, scale=1.0):
    out = torch.zeros((M, C), device=G.DEV)
    G._lib.check(lib.nerf_b200_exp_tile_unpack(G.ptr(img), M, C, scale, G.ptr(out), G.stream()), "tile_unpack")
    torch.cuda.synchronize()
    return out.cpu().numpy()


def f16(x):
    return x.astype(np.float16).astype(np.float64)


@pytest.mark.parametrize("C", [64, 128, 256])
def test_tile_pack_unpack_colsum(G, C):
    lib = G._lib.load()
    x = np.random.default_rng(C).standard_normal((300, C)).astype(np.float32)
    img = _pack(G, lib, x, 4.0)
    back = _unpack(G, lib, img, 300, C, 0.25)
    assert np.array_equal(back, (x * 4).astype(np.float16).astype(np.float32) * 0.25)
    cs = torch.zeros(C, device=G.DEV)
    G._lib.check(lib.nerf_b200_exp_tile_colsum(G.ptr(img), 3, C, 0.25, G.ptr(cs), G.stream()), "tile_colsum")
    torch.cuda.synchronize()
    assert rel_l2(cs.cpu().numpy(), f16(x * 4).sum(0) * 0.25) < 1e-5


@pytest.mark.parametrize("Mc,Nc", [(256, 256), (128, 256), (256, 64), (128, 128)])
def test_wgrad_tiles(G, Mc, Nc):
    """dW = scale * X^T Y over 1000 sample rows (8 tiles, the last one zero-padded), accumulated on top of dW0."""
    lib = G._lib.load()
    rng = np.random.default_rng(Mc + Nc)
    X = rng.standard_normal((1000, Mc)).astype(np.float32)
    Y = rng.standard_normal((1000, Nc)).astype(np.float32)
    dW0 = rng.standard_normal((Mc, Nc)).astype(np.float32)
    xi, yi = _pack(G, lib, X), _pack(G, lib, Y)
    dW = G.dev(dW0.copy())
    G._lib.check(lib.nerf_b200_exp_wgrad_tiles(G.ptr(xi), G.ptr(yi), 8, Mc, Nc, 0.5, G.ptr(dW), Nc, G.stream()), "wgrad_tiles")
    torch.cuda.synchronize()
    ref = dW0 + 0.5 * (f16(X).T @ f16(Y))
    assert rel_l2(dW.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("Kc", [256, 128])
@pytest.mark.parametrize("mask", [False, True])
def test_dgrad_tiles(G, Kc, mask):
    """OUT = relu_mask(H)(X W) per 128-row tile, W [Kc, 256]; 700 rows = 6 tiles on up to 6 CTAs (and on 2 CTAs: 3 tiles each)."""
    lib = G._lib.load()
    rng = np.random.default_rng(Kc + int(mask))
    X = rng.standard_normal((700, Kc)).astype(np.float32)
    W = (rng.standard_normal((Kc, 256)) / np.sqrt(Kc)).astype(np.float32)
    H = np.maximum(rng.standard_normal((700, 256)), 0).astype(np.float32)
    xi, wi, hi = _pack(G, lib, X), _pack(G, lib, W), _pack(G, lib, H)
    oi = torch.zeros(_img_bytes(700, 256), dtype=torch.uint8, device=G.DEV)
    G._lib.check(lib.nerf_b200_exp_dgrad_tiles(G.ptr(xi), G.ptr(wi), G.ptr(hi) if mask else None, 6, Kc, G.ptr(oi), G.stream()), "dgrad_tiles")
    out = _unpack(G, lib, oi, 700, 256)
    ref = f16(X) @ f16(W)
    if mask:
        ref = ref * (f16(H) > 0)
    assert rel_l2(out, ref) < 1e-3          # fp16 rounding of the output image


def test_backward_with_tensor_core_gemms_matches_oracle():
    """NERF_B200_BWD_TC=1 routes the backward's large GEMMs (wgrad / dgrad of the 256- and 128-wide layers) through
    wgrad_tiles_kernel / dgrad_tiles_kernel with fp16 operands and a static loss scale; the forward recompute, the
    ReLU masks and the small GEMMs stay fp32, so the gradients should stay within the exact test's budget.
    Subprocess: the switch is latched at first use."""
    import subprocess, sys, tempfile
    code = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import gpu_common as G
sb = G.synth.ray_batch("lego", 96, seed=4)
pc, pf = G.synth.nerf_state(0), G.synth.nerf_state(1)
nets = [G.make_net(pc), G.make_net(pf)]
target = np.random.default_rng(3).random((96, 3), dtype=np.float32)
G.nb.set_precision("fp32")
rgb, _, _, ex = G.nb.render(400, 400, sb["K"], rays=G.dev(sb["rays"]), ndc=False, near=2., far=6., use_viewdirs=True,
                            network_fn=nets[0], network_fine=nets[1], network_query_fn=G.query_fn(), N_samples=64,
                            N_importance=128, perturb=0., white_bkgd=True, raw_noise_std=0.)
loss = G.nb.img2mse(rgb, G.dev(target)) + G.nb.img2mse(ex["rgb0"], G.dev(target))
loss.backward()
packed = G.O.pack_rays(400, 400, sb["K"], sb["rays"][0], sb["rays"][1], False, 2.0, 6.0, True)
l_ref, gc, gf = G.O.render_rays_grads(packed, pc, pf, 64, 128, target, white_bkgd=True)
errs = []
for net, g in ((nets[0], gc), (nets[1], gf)):
    for name, p in net.named_parameters():
        a, b = p.grad.cpu().numpy().astype(np.float64), g[name].astype(np.float64)
        errs.append(float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)))
np.save(sys.argv[1], np.array(errs))
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "errs.npy")
        subprocess.run([sys.executable, "-c", code, path], check=True, cwd=root, env=dict(os.environ, NERF_B200_BWD_TC="1"), timeout=600)
        errs = np.load(path)
    assert np.median(errs) < 3e-3, errs
    assert errs.max() < 5e-2, errs
