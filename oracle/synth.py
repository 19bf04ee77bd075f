"""Deterministic synthetic inputs for the NeRF hot path  --  TEST INFRASTRUCTURE.

numpy-only so the same bytes are produced in the build container (next to the imported
reference) and on the GPU box (where /root/reference does not exist).
"""
from __future__ import annotations

import numpy as np

LEGO_CAMERA_ANGLE_X = 0.6911112070083618       # dataset constant of nerf_synthetic/lego


def pose_spherical(theta, phi, radius):
    """load_blender.py:11-34 restated in numpy (float32 like torch.Tensor)."""
    t = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], np.float32)
    ph = phi / 180.0 * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(ph), -np.sin(ph), 0], [0, np.sin(ph), np.cos(ph), 0],
                   [0, 0, 0, 1]], np.float32)
    th = theta / 180.0 * np.pi
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0],
                   [0, 0, 0, 1]], np.float32)
    flip = np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32)
    return (flip @ (rt @ (rp @ t))).astype(np.float32)


def intrinsics(H, W, focal):
    """run_nerf.py:615-620."""
    return np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]], np.float32)


def lego_camera(res=400, theta=30.0):
    H = W = res
    focal = 0.5 * W / np.tan(0.5 * LEGO_CAMERA_ANGLE_X)     # load_blender.py:72-73
    return H, W, intrinsics(H, W, focal), pose_spherical(theta, -30.0, 4.0)[:3, :4]


def fern_camera():
    H, W, focal = 378, 504, 407.5658
    c2w = np.concatenate([np.eye(3, dtype=np.float32), np.zeros((3, 1), np.float32)], 1)
    return H, W, intrinsics(H, W, focal), c2w


def camera_rays(H, W, K, c2w):
    """run_nerf_helpers.py:165-172 (get_rays_np)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1).astype(np.float32)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape).astype(np.float32)
    return rays_o, rays_d


def ray_batch(scene="lego", N=4096, seed=0, res=400):
    """-> H, W, K, rays [2,N,3] (o,d), near, far, ndc, white_bkgd for a lego- or fern-shaped batch."""
    if scene == "lego":
        H, W, K, c2w = lego_camera(res)
        near, far, ndc, white = 2.0, 6.0, False, True
    elif scene == "fern":
        H, W, K, c2w = fern_camera()
        near, far, ndc, white = 0.0, 1.0, True, False
    else:
        raise ValueError(scene)
    o, d = camera_rays(H, W, K, c2w)
    rng = np.random.default_rng(seed)
    idx = rng.permutation(H * W)[:N] if N <= H * W else rng.integers(0, H * W, N)
    rays = np.stack([o.reshape(-1, 3)[idx], d.reshape(-1, 3)[idx]], 0).astype(np.float32)
    return dict(H=H, W=W, K=K, c2w=c2w, rays=rays, near=near, far=far, ndc=ndc, white_bkgd=white)


_SHAPES_VIEWDIRS = [("pts_linears.0", 256, 63)] + [(f"pts_linears.{i}", 256, 256) for i in (1, 2, 3, 4)] + \
    [("pts_linears.5", 256, 319), ("pts_linears.6", 256, 256), ("pts_linears.7", 256, 256),
     ("views_linears.0", 128, 283), ("feature_linear", 256, 256), ("alpha_linear", 1, 256),
     ("rgb_linear", 3, 128)]


def nerf_state(seed=0, sharpen=False, D=8, W=256, input_ch=63, input_ch_views=27, skips=(4,)):
    """Deterministic NeRF parameters with nn.Linear's default init *distribution*
    (U(-1/sqrt(in), 1/sqrt(in)) for weight and bias), keyed like the reference state_dict
    (run_nerf_helpers.py:79-94), in the reference's parameter order.

    sharpen=True applies SURVEY 8d(ii): alpha_linear.weight*30, bias+0.5, rgb_linear.weight*5.
    """
    rng = np.random.default_rng(1000 + seed)
    shapes = [("pts_linears.0", W, input_ch)]
    for i in range(D - 1):
        shapes.append((f"pts_linears.{i + 1}", W, W + input_ch if i in skips else W))
    shapes += [("views_linears.0", W // 2, input_ch_views + W), ("feature_linear", W, W),
               ("alpha_linear", 1, W), ("rgb_linear", 3, W // 2)]
    p = {}
    for name, out_f, in_f in shapes:
        b = 1.0 / np.sqrt(in_f)
        p[name + ".weight"] = rng.uniform(-b, b, (out_f, in_f)).astype(np.float32)
        p[name + ".bias"] = rng.uniform(-b, b, (out_f,)).astype(np.float32)
    if sharpen:
        p["alpha_linear.weight"] = p["alpha_linear.weight"] * np.float32(30)
        p["alpha_linear.bias"] = p["alpha_linear.bias"] + np.float32(0.5)
        p["rgb_linear.weight"] = p["rgb_linear.weight"] * np.float32(5)
    return p


def rng_draws(N, N_samples, N_importance, seed=0):
    """Injected RNG draws: t_rand [N,N_samples], u [N,N_importance], both U[0,1) float32."""
    rng = np.random.default_rng(77 + seed)
    t_rand = rng.random((N, N_samples), dtype=np.float32)
    u = rng.random((N, max(N_importance, 1)), dtype=np.float32)[:, :N_importance]
    return t_rand, u
