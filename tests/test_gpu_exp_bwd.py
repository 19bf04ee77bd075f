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
