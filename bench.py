#!/usr/bin/env python
"""bench.py -- rays/sec of the NeRF ray-marching hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path (render_rays: 64 coarse + 128 fine samples, 8x256 MLP +
128-wide view head) over one 4096-ray batch of synthetic lego-shaped rays (BASELINE configs[1]).
With N GPUs every rank renders its own 4096-ray batch (rays shard with no data-path collective:
"scaling": "weak"); `value` = rays of all ranks / max-over-ranks device time.

JSON keys: see the build spec (value, e2e, roofline, cpu_baseline, clocks, gpu_launches ...).
`--impl reference` times the reference's op chain on the host cores (oracle/torch_ref.py: stock torch
fp32 CPU ops at the reference's granularity, all threads; pinned to the reference-generated golden
vectors) on a bounded sample of the same workload; the reference itself is Python on torch and
/root/reference does not exist on the GPU box.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_SAMPLES, N_IMPORTANCE = 4096, 64, 128
FLOP_PER_RAY_FWD = 303_824_896          # 593 408 MAC x 2 x 256 evaluations
WORKLOAD = "lego 400x400 synthetic rays, N_rand=4096, N_samples=64 + N_importance=128, D=8 W=256 use_viewdirs, forward render_rays"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            pk = json.load(f)
        return float(pk["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst; fp16 == bf16 rate)", pk.get("bf16_tflops_sustained")
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)", None


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on the host cores
# --------------------------------------------------------------------------------------------

def _cpu_limit():
    """Host threads this process may use: scheduler affinity capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:
            q, per = f.read().split()[:2]
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


def cpu_port_rays_per_s(n_rays, reps, warmup, budget_s=40.0):
    """The reference's op chain (stock torch ops at the reference's granularity, oracle/torch_ref.py -- pinned to the
    reference-generated golden vectors) on the host cores: what run_nerf.py's render() does on a CPU.  The thread
    count is calibrated (more threads than the container really owns slows torch's CPU ops down badly: 128 threads
    on the GPU box gave 41 rays/s), the measurement is time-bounded."""
    import torch
    from oracle import synth, torch_ref as T
    sb = synth.ray_batch("lego", n_rays, seed=0)
    rays = torch.from_numpy(sb["rays"])
    sd = [{k: torch.from_numpy(v) for k, v in synth.nerf_state(s).items()} for s in (0, 1)]

    def run(n):
        t0 = time.perf_counter()
        with torch.no_grad():
            T.render(rays[0, :n], rays[1, :n], sd[0], sd[1], 2.0, 6.0, S=N_SAMPLES, n_imp=N_IMPORTANCE, white_bkgd=True)
        return time.perf_counter() - t0

    limit = _cpu_limit()
    best_t, best = None, None
    for th in sorted({c for c in (4, 8, 16, 32, 64, limit) if c <= limit}):
        torch.set_num_threads(th)
        run(64)
        dt = run(128)
        if best is None or dt < best:
            best_t, best = th, dt
    torch.set_num_threads(best_t)
    ts, t_start = [], time.perf_counter()
    for i in range(warmup + reps):
        dt = run(n_rays)
        if i >= warmup:
            ts.append(dt)
        if ts and time.perf_counter() - t_start > budget_s:
            break
    return n_rays / (sum(ts) / len(ts)), sum(ts) / len(ts), best_t


def reference_arm(args, rank, world):
    if rank != 0:
        return
    n = 1024                                          # bounded sample of the 4096-ray batch per step
    steps, warmup = max(1, min(args.steps, 10)), max(1, min(args.warmup, 2))
    rps, sec, threads = cpu_port_rays_per_s(n, steps, warmup)
    line = {"impl": "reference", "metric": "rays/sec", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3 * (N_RAYS / n), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n} of the 4096 rays per step"},
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                             "sample": f"{n} rays x (64+128) samples per step, torch fp32 CPU ops (the reference's op chain), {threads} threads (calibrated; {os.cpu_count()} logical CPUs visible)"},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks sampling
# --------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-rays", type=int, default=1024, help="rays in the cpu_baseline sample (rank 0, N=1)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    assert torch.cuda.is_available(), "bench.py needs a GPU (use --impl reference for the CPU arm)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import nerf_pytorch_b200 as nb
    from nerf_pytorch_b200 import _lib
    from nerf_pytorch_b200.api import _QueryFn
    from oracle import synth
    lib = _lib.load()

    # ---- workload: weights (same on all ranks), rays (per-rank batch) ----
    nets = []
    for seed in (0, 1):
        m = nb.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, output_ch=5, skips=[4], use_viewdirs=True)
        m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.nerf_state(seed).items()})
        nets.append(m.to(dev))
    e, _ = nb.get_embedder(10, 0)
    ed, _ = nb.get_embedder(4, 0)
    q = _QueryFn(e, ed, 65536, 10, 4, 0)
    sb = synth.ray_batch("lego", N_RAYS, seed=rank)
    kw = dict(ndc=False, near=2., far=6., use_viewdirs=True, network_fn=nets[0], network_fine=nets[1], network_query_fn=q,
              N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0., white_bkgd=True, raw_noise_std=0.)
    rays_host = torch.from_numpy(sb["rays"]).pin_memory()
    rays_dev = rays_host.to(dev)
    out_host = torch.empty((N_RAYS, 5), dtype=torch.float32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)            # > 126 MB L2

    def step_resident():
        with torch.no_grad():
            return nb.render(sb["H"], sb["W"], sb["K"], chunk=32768, rays=rays_dev, **kw)

    graphed = nb.GraphedRender(sb["H"], sb["W"], sb["K"], N_RAYS, chunk=32768, **kw)

    def step_e2e():
        # the user-facing call: pinned host rays in, pinned host [rgb, disp, acc] out (H2D + graph replay + D2H + sync)
        return graphed(rays_host)

    def timed(fn, steps, warmup, use_events=True):
        for _ in range(warmup):
            flush.zero_(); fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs, wall = [], 0.0
        for _ in range(steps):
            flush.zero_()                                   # L2 flush between timed iterations (not timed)
            if use_events:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                evs.append((a, b))
            else:
                torch.cuda.synchronize()
                t0 = time.perf_counter(); fn(); wall += time.perf_counter() - t0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs) if use_events else wall * 1e3
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    sampler = ClockSampler(local_rank)
    sampler.start()
    step_resident(); torch.cuda.synchronize()
    l0 = nb.launch_count(); step_resident(); launches_per_step = nb.launch_count() - l0
    lib.nerf_b200_timing_enable(1)
    ms_res = timed(step_resident, args.steps, args.warmup)
    kms, kn, kfl = C.c_double(), C.c_int64(), C.c_double()
    # the accumulators also hold the warm-up launches: per-launch averages are what we report
    lib.nerf_b200_timing_read(C.byref(kms), C.byref(kn), C.byref(kfl))
    lib.nerf_b200_timing_enable(0)
    launches = launches_per_step * args.steps
    ms_e2e = timed(step_e2e, args.steps, args.warmup, use_events=False)     # host-side clock: copies + sync included
    clocks = sampler.stop()

    value = N_RAYS * world / (ms_res * 1e-3)
    e2e = N_RAYS * world / (ms_e2e * 1e-3)
    peak, peak_src, peak_sus = load_peaks()
    achieved = (kfl.value / (kms.value * 1e-3)) / 1e12 if kms.value > 0 else 0.0
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "march_tc_traffic.json")
    if os.path.isfile(tfile):
        with open(tfile) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    line = {
        "metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_gpu_per_step": N_RAYS, "operands": "fp16 x fp16 -> fp32 accumulate (tcgen05 kind::f16)",
                   "l2": "flushed between timed iterations (512 MiB memset, outside the timed events)",
                   "parallelism": f"ray-parallel x{world}, no data-path collective"},
        "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": int(rays_host.numel() * 4),
                "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": ms_e2e,
                "api": "nerf_pytorch_b200.GraphedRender(...)(rays_host): H2D copy, CUDA-graph replay of render(), D2H copy of rgb/disp/acc, stream sync"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": ("march_tc_kernel" if os.environ.get("NERF_B200_PAIR", "1")[:1] == "0" else "march_tc2_kernel (cta_group::2 pair)") + " (coarse + fine launches)", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_source": peak_src,
                     # the same achieved rate against cuBLAS' back-to-back (power-limited) bf16 rate: the timed steps run
                     # back to back, so this is the ceiling the chip actually sustains
                     "peak_sustained": peak_sus, "frac_sustained": (achieved / peak_sus) if peak_sus else None,
                     "kernel_ms_per_step": kms.value / max(1, kn.value) * 2,
                     "launches_timed": int(kn.value)},
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu:
        rps, sec, threads = cpu_port_rays_per_s(args.cpu_rays, 5, 1)
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_rays} rays x (64+128) samples, <= 5 reps after 1 warm-up, torch fp32 CPU ops (the reference's op chain), {threads} threads (calibrated; {os.cpu_count()} logical CPUs visible), {sec:.2f} s/rep"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
