"""Pin the quantisation-aware oracle (oracle/tc_emul.py) on the CPU: its hand-written backward must be the autograd
gradient of its own forward when every fp16 rounding is a straight-through estimator, up to the rounding of the
activation gradients themselves (which autograd does not do)."""
import numpy as np
import torch

from oracle import synth, tc_emul as E, torch_ref as T


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float64)

    @staticmethod
    def backward(ctx, g):
        return g


def _forward_ste(sd, pts, vd, D=8, skip=4, ic=63):
    r = _RoundSTE.apply
    enc16 = r(T.embed(pts.float(), 10).double())
    encv = T.embed(vd.float(), 4).double()
    h16 = None
    for l in range(D):
        W, b = r(sd[f"pts_linears.{l}.weight"]), sd[f"pts_linears.{l}.bias"]
        if l == 0:
            pre = enc16 @ W.t() + b
        elif l == skip + 1:
            pre = enc16 @ W[:, :ic].t() + h16 @ W[:, ic:].t() + b
        else:
            pre = h16 @ W.t() + b
        h32 = torch.relu(pre)
        h16 = r(h32)
    sigma = h32 @ sd["alpha_linear.weight"].t() + sd["alpha_linear.bias"]
    feat16 = r(h16 @ r(sd["feature_linear.weight"]).t() + sd["feature_linear.bias"])
    Wv = sd["views_linears.0.weight"]
    pre_v = feat16 @ r(Wv)[:, :256].t() + encv @ Wv[:, 256:].t() + sd["views_linears.0.bias"]
    hv = torch.relu(pre_v)
    rgb = hv @ sd["rgb_linear.weight"].t() + sd["rgb_linear.bias"]
    return torch.cat([rgb, sigma], -1)


def test_emulated_backward_is_the_gradient_of_the_emulated_forward():
    rng = np.random.default_rng(0)
    M, n_rays, S = 96, 6, 16
    pts = torch.from_numpy(rng.uniform(-2, 2, (M, 3)).astype(np.float32))
    dirs = rng.standard_normal((n_rays, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
    vd = torch.from_numpy(np.repeat(dirs, S, 0))
    state = synth.nerf_state(3)
    sd64 = {k: torch.from_numpy(v).double() for k, v in state.items()}
    em = E.forward({k: torch.from_numpy(v) for k, v in state.items()}, pts, vd)
    # autograd through the STE forward
    sd_g = {k: v.clone().requires_grad_(True) for k, v in sd64.items()}
    raw = _forward_ste(sd_g, pts, vd)
    assert torch.allclose(raw, em["raw"], rtol=1e-9, atol=1e-9)
    d_raw = torch.from_numpy(rng.standard_normal((M, 4))).double() * 1e-3
    (raw * d_raw).sum().backward()
    masks = {"h": [(p > 0).double() for p in em["pre"]], "hv": (em["pre_v"] > 0).double()}
    ray_of_row = torch.arange(M) // S
    scale = 2.0 ** 16
    g, st = E.backward({k: torch.from_numpy(v) for k, v in state.items()}, em, masks, d_raw, scale, ray_of_row, n_rays)
    for name, ref in sd_g.items():
        a, b = g[name].reshape(-1), ref.grad.reshape(-1)
        err = float((a - b).norm() / b.norm())
        # the emulation rounds every activation gradient to fp16 (11 bits) once per layer: <= ~1e-3 after nine layers
        assert err < 2e-3, (name, err)
    # the loss scale is a power of two: apart from values that leave / enter the fp16 subnormal range the result does not depend on it
    g2, _ = E.backward({k: torch.from_numpy(v) for k, v in state.items()}, em, masks, d_raw, scale * 4, ray_of_row, n_rays)
    for name in g:
        assert float((g[name] - g2[name]).norm() / g2[name].norm()) < 1e-5, name
