"""One optimisation step of the reference's training loop as ONE replayed CUDA graph.

`run_nerf.train()` does, per iteration (run_nerf.py:757-784): render the ray batch with `retraw=True`, two `img2mse`
terms, `loss.backward()`, `optimizer.step()` (Adam over 48 tensors), then the exponential learning-rate decay on the
host.  Through the drop-in that is ~40 library launches plus autograd and optimizer bookkeeping in Python; after the
fusion the GPU work is a few milliseconds, so the host side matters.  `FusedTrainStep` runs the same arithmetic with
no autograd graph and no per-tensor optimizer:

    pack weights -> pack rays -> z sampling -> fused pass (coarse, training mode) -> resampling -> fused pass (fine)
    -> fused MSE + gradient seed (both terms) -> tensor-core backward of both passes into ONE flat gradient buffer
    -> [one all-reduce of that buffer when data parallel] -> flat Adam with the decayed rate computed on the device

captured once for a fixed ray count and replayed.  The parameters stay the SAME `nn.Parameter` objects (re-pointed at
slices of one flat buffer), so `state_dict()` / checkpoints / `render()` keep working; `attach_optimizer()` exposes the
moments to a `torch.optim.Adam` so that the reference's checkpoint format (run_nerf.py:792-800) is unchanged.
No reference counterpart; nothing here is needed by the drop-in path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from . import api
from ._lib import NerfPassOut, NerfTrainSave, check


class FusedTrainStep:
    def __init__(self, H, W, K, n_rays, render_kwargs, lrate=5e-4, lrate_decay=250, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, use_graph=True, data_parallel=True):
        """render_kwargs: the `render_kwargs_train` dict of create_nerf() plus near / far (run_nerf.py:643-648).
        lrate_decay is in units of 1000 steps like the reference's --lrate_decay (run_nerf.py:779-780)."""
        import torch.distributed as dist
        self.lib = _lib.load()
        kw = dict(render_kwargs)
        self.net_c, self.net_f = kw["network_fn"], kw.get("network_fine")
        if not isinstance(self.net_c, api.NeRF) or not self.net_c.use_viewdirs:
            raise RuntimeError("FusedTrainStep needs nerf_b200.NeRF networks with use_viewdirs=True")
        if api.get_precision() != "tc_fp16":
            raise RuntimeError("FusedTrainStep runs the tensor-core path (set_precision('tc_fp16'))")
        self.dev = self.net_c._check_device()
        self.H, self.W, self.K, self.N = int(H), int(W), K, int(n_rays)
        self.Sc, self.Ni = int(kw["N_samples"]), int(kw.get("N_importance", 0))
        if self.Ni > 0 and self.net_f is None:
            self.net_f = self.net_c
        q = kw.get("network_query_fn")
        self.cfgd = dict(N_samples=self.Sc, N_importance=self.Ni, multires=int(getattr(q, "multires", 10)),
                         multires_views=int(getattr(q, "multires_views", 4)), lindisp=bool(kw.get("lindisp", False)),
                         perturb=float(kw.get("perturb", 0.)), white_bkgd=bool(kw.get("white_bkgd", False)), retraw=True, want_grad=True)
        self.noise_std = float(kw.get("raw_noise_std", 0.))
        self.ndc, self.near, self.far = int(bool(kw.get("ndc", True))), float(kw.get("near", 0.)), float(kw.get("far", 1.))
        self.lr0, self.decay_rate, self.decay_steps = float(lrate), 0.1, float(lrate_decay) * 1000.0
        self.b1, self.b2, self.eps = float(betas[0]), float(betas[1]), float(eps)
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (data_parallel and dist.is_available() and dist.is_initialized()) else 1
        # ---- one flat buffer for parameters, gradients and moments; the Parameters become views of it ----
        nets = [self.net_c] + ([self.net_f] if (self.net_f is not None and self.net_f is not self.net_c) else [])
        self.params = [p for n in nets for p in n.parameters()]
        total = sum(p.numel() for p in self.params)
        f32 = dict(device=self.dev, dtype=torch.float32)
        self.flat_p = torch.empty(total, **f32)
        self.flat_g = torch.zeros(total, **f32)
        self.m, self.v = torch.zeros(total, **f32), torch.zeros(total, **f32)
        self.state = torch.zeros(4, **f32)          # [0] loss of the last step, [1] steps taken, [2] rate of the last step
        off, self._gviews = 0, {}
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                self.flat_p[off:off + n].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + n].view_as(p)
                p.grad = self.flat_g[off:off + n].view_as(p)
                off += n
        self._gs = []
        for n in nets:
            self._gs.append(n.grad_struct({k: p.grad for k, p in n.named_parameters()}))
        self._nets = nets
        # ---- fixed buffers of the step ----
        N, Sc, Sf = self.N, self.Sc, self.Sc + self.Ni
        self.rays = torch.zeros((2, N, 3), **f32)
        self.rays[1, :, 2] = -1.0
        self.target = torch.zeros((N, 3), **f32)
        self.packed_rays = torch.empty((N, 11), **f32)
        self.t_rand = torch.zeros((N, Sc), **f32) if self.cfgd["perturb"] > 0 else None
        self.u_rand = torch.zeros((N, self.Ni), **f32) if (self.cfgd["perturb"] > 0 and self.Ni > 0) else None
        self.noise0 = torch.zeros((N, Sc), **f32) if self.noise_std > 0 else None
        self.noise1 = torch.zeros((N, Sf), **f32) if (self.noise_std > 0 and self.Ni > 0) else None
        self.t_vals = torch.linspace(0., 1., steps=Sc, device=self.dev)
        self.u_det = torch.linspace(0., 1., steps=self.Ni, device=self.dev) if self.Ni > 0 else None
        self.z_c, self.z_f, self.z_std = torch.empty((N, Sc), **f32), torch.empty((N, max(Sf, 1)), **f32), torch.empty((N,), **f32)
        o = {k: torch.empty(s, **f32) for k, s in (("rgb0", (N, 3)), ("disp0", (N,)), ("acc0", (N,)), ("w0", (N, Sc)), ("raw0", (N, Sc, 4)),
                                                   ("rgb", (N, 3)), ("disp", (N,)), ("acc", (N,)), ("raw", (N, Sf, 4)),
                                                   ("g_rgb", (N, 3)), ("g_rgb0", (N, 3)))}
        self.out = o
        self._out_c = NerfPassOut(api._ptr(o["rgb0"]), api._ptr(o["disp0"]), api._ptr(o["acc0"]), C.c_void_p(0), api._ptr(o["w0"]), api._ptr(o["raw0"]))
        self._out_f = NerfPassOut(api._ptr(o["rgb"]), api._ptr(o["disp"]), api._ptr(o["acc"]), C.c_void_p(0), C.c_void_p(0), api._ptr(o["raw"]))
        self._np_c, self._np_f = self.net_c.net_params(), (self.net_f.net_params() if self.Ni > 0 else None)
        lib = self.lib
        self._pk_c = torch.empty(lib.nerf_b200_packed_bytes(C.byref(self._np_c)), dtype=torch.uint8, device=self.dev)
        self._pk_f = self._pk_c if (self.Ni == 0 or self.net_f is self.net_c) else \
            torch.empty(lib.nerf_b200_packed_bytes(C.byref(self._np_f)), dtype=torch.uint8, device=self.dev)
        with api._on(self.rays):
            self._sv_c, self._act_c, self._mask_c = api._train_save(lib, N, Sc, self._np_c, self.dev)
            if self.Ni > 0:
                self._sv_f, self._act_f, self._mask_f = api._train_save(lib, N, Sf, self._np_f, self.dev)
            self._ws = torch.empty(lib.nerf_b200_march_workspace_bytes(N, Sf), dtype=torch.uint8, device=self.dev)
            bw = lib.nerf_b200_render_rays_bwd_tc_workspace_bytes(N, Sc, C.byref(self._np_c), Sf if self.Ni > 0 else 0,
                                                                  C.byref(self._np_f) if self.Ni > 0 else None)
            self._bws = torch.empty(bw + 1024, dtype=torch.uint8, device=self.dev)
        o = self.out
        self._bp_c = _lib.NerfBwdPass(api._ptr(self.z_c), api._ptr(self.noise0), Sc, C.pointer(self._np_c), api._ptr(self._pk_c), api._ptr(o["raw0"]),
                                      C.pointer(self._sv_c), api._ptr(o["g_rgb0"]), C.pointer(self._gs[0]))
        self._bp_f = _lib.NerfBwdPass(api._ptr(self.z_f), api._ptr(self.noise1), Sf, C.pointer(self._np_f), api._ptr(self._pk_f), api._ptr(o["raw"]),
                                      C.pointer(self._sv_f), api._ptr(o["g_rgb"]), C.pointer(self._gs[-1])) if self.Ni > 0 else None
        self._cam = api._camera(self.H, self.W, self.K)
        self._cfg = api._cfg_struct(self.cfgd, 11)
        self.host_loss = torch.zeros(4, dtype=torch.float32, device="cpu").pin_memory()
        # ---- capture ----
        self.graph = self.graph_adam = None
        if use_graph:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            snap = (self.flat_p.clone(), self.m.clone(), self.v.clone(), self.state.clone())
            with torch.cuda.stream(side):
                for _ in range(2):                      # opt-ins, caches, lazy initialisations outside the capture
                    self._pack_rays(); self._fwd_bwd(); self._adam()
            torch.cuda.current_stream(self.dev).wait_stream(side)
            torch.cuda.synchronize(self.dev)
            with torch.no_grad():                       # the warm-up steps must not count: restore parameters and moments
                self.flat_p.copy_(snap[0]); self.m.copy_(snap[1]); self.v.copy_(snap[2]); self.state.copy_(snap[3])
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._fwd_bwd()
                if self.world == 1:
                    self._adam()
            if self.world > 1:
                self.graph_adam = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_adam):
                    self._adam()

    # ---- the step, as library calls on the current stream ----
    def _pack_rays(self, c2w=None, pixel_index=None):
        """The ray batch -> packed [N,11] rows.  Outside the graph (one launch): the camera travels in the kernel arguments, so a
        per-step pose could not be replayed.  Given rays (self.rays) or -- SURVEY 8f rank 3 -- pixels of one posed image."""
        with api._on(self.rays):
            if pixel_index is None:
                check(self.lib.nerf_b200_pack_rays(api._ptr(self.rays[0]), api._ptr(self.rays[1]), None, C.byref(self._cam), self.N, 0, self.ndc, self.near,
                                                   self.far, 1, api._ptr(self.packed_rays), api._stream(self.rays)), "pack_rays")
            else:
                cam = api._camera(self.H, self.W, self.K, c2w)
                check(self.lib.nerf_b200_pack_rays_pixels(C.byref(cam), api._ptr(pixel_index), self.N, self.ndc, self.near, self.far, 1,
                                                          api._ptr(self.packed_rays), api._stream(self.rays)), "pack_rays_pixels")

    def _fwd_bwd(self):
        lib, N, cfg = self.lib, self.N, self._cfg
        st = api._stream(self.rays)
        with api._on(self.rays):
            check(lib.nerf_b200_pack_weights(C.byref(self._np_c), api._ptr(self._pk_c), self._pk_c.numel(), st), "pack_weights")
            if self._pk_f is not self._pk_c:
                check(lib.nerf_b200_pack_weights(C.byref(self._np_f), api._ptr(self._pk_f), self._pk_f.numel(), st), "pack_weights")
            if self.t_rand is not None:                                 # (the reference's draw order)
                self.t_rand.uniform_()                                  # run_nerf.py:371
            if self.noise0 is not None:
                self.noise0.normal_().mul_(self.noise_std)              # run_nerf.py:285
            if self.u_rand is not None:
                self.u_rand.uniform_()                                  # run_nerf_helpers.py:208
            if self.noise1 is not None:
                self.noise1.normal_().mul_(self.noise_std)
            fine = self.Ni > 0
            check(lib.nerf_b200_render_rays_fwd_train(
                api._ptr(self.packed_rays), N, C.byref(cfg), C.byref(self._np_c), api._ptr(self._pk_c),
                C.byref(self._np_f) if fine else None, api._ptr(self._pk_f) if fine else None,
                api._ptr(self.t_vals), api._ptr(self.u_det), api._ptr(self.t_rand), api._ptr(self.u_rand), api._ptr(self.noise0), api._ptr(self.noise1),
                api._ptr(self.z_c), C.byref(self._out_c), api._ptr(self.z_f) if fine else None, api._ptr(self.z_std) if fine else None,
                C.byref(self._out_f) if fine else None, api._ptr(self._ws), self._ws.numel(),
                C.byref(self._sv_c), C.byref(self._sv_f) if fine else None, st), "render_rays_fwd_train")
            # loss = img2mse(rgb, target) + img2mse(rgb0, target)  (run_nerf.py:764-772) and its gradient seeds
            self.state[0:1].zero_()
            self.flat_g.zero_()
            gscale = 1.0 / self.world
            o = self.out
            if fine:
                check(lib.nerf_b200_mse_seed(api._ptr(o["rgb"]), api._ptr(self.target), N, gscale, api._ptr(o["g_rgb"]), api._ptr(self.state), st), "mse_seed")
            check(lib.nerf_b200_mse_seed(api._ptr(o["rgb0"]), api._ptr(self.target), N, gscale, api._ptr(o["g_rgb0"]), api._ptr(self.state), st), "mse_seed")
            # backward of both passes into the flat gradient buffer (one call: the passes are scheduled side by side)
            check(lib.nerf_b200_render_rays_bwd_tc(api._ptr(self.packed_rays), N, C.byref(cfg), C.byref(self._bp_c),
                                                   C.byref(self._bp_f) if fine else None, api._ptr(self._bws), self._bws.numel(), st), "render_rays_bwd_tc")

    def _adam(self):
        with api._on(self.rays):
            check(self.lib.nerf_b200_adam_step(api._ptr(self.flat_p), api._ptr(self.flat_g), api._ptr(self.m), api._ptr(self.v), self.flat_p.numel(),
                                               api._ptr(self.state), self.lr0, self.decay_rate, self.decay_steps, self.b1, self.b2, self.eps, 1.0,
                                               api._stream(self.rays)), "adam_step")

    # ---- public ----
    def step_device(self, c2w=None, pixel_index=None):
        """One step on the rays / target already in `self.rays` ([2,N,3]) and `self.target` ([N,3]); no host sync.
        With `c2w` [3,4] and `pixel_index` (int64 CUDA tensor [N] of row-major pixel ids) the rays are generated on the device
        for those pixels of that pose instead (run_nerf.py:728-757 without get_rays' [H,W,3] tensors)."""
        import torch.distributed as dist
        self._pack_rays(c2w, pixel_index)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._fwd_bwd()
        if self.world > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)     # ONE collective: every gradient, 4.77 MB
            if self.graph_adam is not None:
                self.graph_adam.replay()
            else:
                self._adam()
        elif self.graph is None:
            self._adam()
        for n in self._nets:
            n.invalidate_pack()                         # parameters changed behind autograd's version counters

    def __call__(self, rays_host, target_host):
        """rays_host [2,N,3], target_host [N,3] (pinned host memory makes the copies asynchronous) -> loss (python float).
        Copies the batch in, runs the step, copies the loss out and synchronises the stream."""
        self.rays.copy_(rays_host, non_blocking=True)
        self.target.copy_(target_host, non_blocking=True)
        self.step_device()
        self.host_loss.copy_(self.state, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        return float(self.host_loss[0])

    @property
    def global_step(self) -> int:
        return int(self.state[1].item())

    def attach_optimizer(self, optimizer: torch.optim.Adam):
        """Make `optimizer`'s per-parameter state alias this step's flat moments, so optimizer.state_dict() (the
        reference's checkpoint content, run_nerf.py:796-799) reflects the fused steps."""
        off = 0
        step = self.state[1].detach().clone().cpu()
        for p in self.params:
            n = p.numel()
            optimizer.state[p] = {"step": step.clone(), "exp_avg": self.m[off:off + n].view_as(p), "exp_avg_sq": self.v[off:off + n].view_as(p)}
            off += n
        t = float(self.state[1].item())
        for g in optimizer.param_groups:                # the rate the NEXT step will use, as run_nerf.py:779-783 leaves it
            g["lr"] = self.lr0 * self.decay_rate ** ((t - 1.0) / self.decay_steps) if (t >= 1.0 and self.decay_steps > 0) else self.lr0
        return optimizer
