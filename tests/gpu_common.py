"""Shared helpers for the -m gpu parity tests (they call the CUDA library through the C ABI via the
Python host mirror and compare with the numpy oracle / the committed golden fixtures)."""
import ctypes as C

import numpy as np
import torch

import nerf_pytorch_b200 as nb
from nerf_pytorch_b200 import _lib
from nerf_pytorch_b200.api import _QueryFn
from oracle import nerf_oracle as O
from oracle import synth

DEV = torch.device("cuda:0")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def make_net(state, use_viewdirs=True, output_ch=5, D=8, skips=(4,)):
    m = nb.NeRF(D=D, W=256, input_ch=63, input_ch_views=27 if use_viewdirs else 0, output_ch=output_ch,
                skips=list(skips), use_viewdirs=use_viewdirs)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
    return m.to(DEV)


def query_fn():
    e, _ = nb.get_embedder(10, 0)
    ed, _ = nb.get_embedder(4, 0)
    return _QueryFn(e, ed, 65536, 10, 4, 0)


def packed_rays(fx):
    rays = fx["rays"]
    return O.pack_rays(int(fx["H"]), int(fx["W"]), fx["K"], rays[0], rays[1], bool(fx["ndc"]),
                       float(fx["near"]), float(fx["far"]), True)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
