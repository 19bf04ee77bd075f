"""on-device ray batching and image-tile output around the hot path (run_nerf.py:677-757, :151-169)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_common
    return gpu_common


def test_pack_rays_pixels_equals_gathering_from_full_image_rays(G):
    """rays of an arbitrary pixel list generated on the device == packing the full image and indexing (the reference's
    get_rays + meshgrid + np.random.choice + gather, run_nerf.py:728-757)."""
    from nerf_pytorch_b200 import api
    lib = G._lib.load()
    H, W, K, c2w = G.synth.lego_camera(60)
    cam = api._camera(H, W, K, c2w)
    full = torch.empty((H * W, 11), device=G.DEV)
    G._lib.check(lib.nerf_b200_pack_rays(None, None, None, C.byref(cam), H * W, 0, 0, 2.0, 6.0, 1, G.ptr(full), G.stream()), "pack_rays")
    pix = torch.randperm(H * W, device=G.DEV)[:777]
    out = torch.empty((777, 11), device=G.DEV)
    G._lib.check(lib.nerf_b200_pack_rays_pixels(C.byref(cam), G.ptr(pix), 777, 0, 2.0, 6.0, 1, G.ptr(out), G.stream()), "pack_rays_pixels")
    torch.cuda.synchronize()
    assert torch.equal(out, full[pix])
    o, d = G.synth.camera_rays(H, W, K, c2w)
    assert np.allclose(out[:, 3:6].cpu().numpy(), d.reshape(-1, 3)[pix.cpu().numpy()], rtol=0, atol=1e-6)


def test_render_to8b_frame(G):
    nets = [G.make_net(G.synth.nerf_state(0)), G.make_net(G.synth.nerf_state(1))]
    H, W, K, c2w = G.synth.lego_camera(48)
    kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=G.query_fn(),
              N_samples=64, N_importance=128, perturb=0., white_bkgd=True, raw_noise_std=0.)
    rgb8, disp = G.nb.render_to8b(H, W, K, c2w, chunk=1000, **kw)           # ragged chunks
    with torch.no_grad():
        rgb, disp_ref, _, _ = G.nb.render(H, W, K, chunk=32768, c2w=torch.from_numpy(c2w), **kw)
    torch.cuda.synchronize()
    assert np.array_equal(rgb8.numpy(), G.nb.to8b(rgb.cpu().numpy()))       # run_nerf_helpers.py:11
    assert torch.equal(torch.nan_to_num(disp), torch.nan_to_num(disp_ref))


def test_device_ray_batcher_and_pixel_train_step(G):
    from nerf_pytorch_b200.trainer import FusedTrainStep
    H, W, K, _ = G.synth.lego_camera(40)
    n_img, B = 5, 256
    rng = np.random.default_rng(0)
    images = rng.random((n_img, H, W, 3), dtype=np.float32)
    poses = np.stack([G.synth.pose_spherical(30.0 * i, -30.0, 4.0) for i in range(n_img)], 0)
    # mode 1: shuffled rays_rgb on the device
    rays = np.stack([np.stack(G.synth.camera_rays(H, W, K, p[:3, :4]), 0) for p in poses], 0)                 # [N, 2, H, W, 3]
    rays_rgb = np.concatenate([rays, images[:, None]], 1).transpose(0, 2, 3, 1, 4).reshape(-1, 3, 3).astype(np.float32)
    b1 = G.nb.DeviceRayBatcher(images, poses, H, W, B, rays_rgb=rays_rgb)
    seen = 0
    for _ in range(2 * (rays_rgb.shape[0] // B) + 3):                        # crosses two epoch boundaries
        br, tg = b1.next()
        assert br.shape == (2, min(B, br.shape[1]), 3) and tg.shape[1] == 3 and br.is_cuda
        seen += br.shape[1]
    assert seen > 2 * rays_rgb.shape[0] - 2 * B
    # mode 2: per-image random pixels, rays generated on the device inside the train step
    b2 = G.nb.DeviceRayBatcher(images, poses, H, W, B, i_train=[0, 1, 2, 3], precrop_iters=2, precrop_frac=0.5)
    nets = [G.make_net(G.synth.nerf_state(0)), G.make_net(G.synth.nerf_state(1))]
    kw = dict(network_fn=nets[0], network_fine=nets[1], network_query_fn=G.query_fn(), N_samples=64, N_importance=128, perturb=1.,
              white_bkgd=True, raw_noise_std=0., use_viewdirs=True, ndc=False, lindisp=False, near=2., far=6.)
    tr = FusedTrainStep(H, W, K, B, kw, lrate=5e-4)
    losses = []
    for it in range(6):
        c2w, pix, target = b2.next()
        if it < 2:                                                           # centre crop (run_nerf.py:731-739)
            jj, ii = (pix // W).cpu().numpy(), (pix % W).cpu().numpy()
            assert jj.min() >= H // 2 - H // 4 and jj.max() < H // 2 + H // 4 and ii.min() >= W // 2 - W // 4 and ii.max() < W // 2 + W // 4
        else:
            assert len(torch.unique(pix)) == B                               # without replacement (:741)
        tr.target.copy_(target)
        tr.step_device(c2w=c2w, pixel_index=pix)
        losses.append(float(tr.state[0].item()))
        # the rays the step used are the image's rays at those pixels
        o, d = G.synth.camera_rays(H, W, K, c2w.numpy())
        assert np.allclose(tr.packed_rays[:, 3:6].cpu().numpy(), d.reshape(-1, 3)[pix.cpu().numpy()], atol=1e-6)
    assert all(np.isfinite(losses)) and tr.global_step == 6
