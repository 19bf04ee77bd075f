"""Print the numbers profiles/r02_summary.md quotes, from the committed captures (no GPU needed):
the bench lines (1 / 2 / 8 GPUs), the launch list of one fused training step, the ncu --set full metrics of the training
kernels, and the measured parity figures.

    python tools/summarize_profiles_r02.py [profiles_dir]
"""
import collections, csv, json, os, sys

d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")


def line(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


b = line(os.path.join(d, "r02_bench_1gpu.json"))
rf, t = b["roofline"], b["train"]
print(f"== 1 GPU: forward {b['value']/1e6:.2f} M rays/s ({b['ms_per_step']:.3f} ms/step), e2e graph {b['e2e']['value']/1e6:.2f} M, eager {b['e2e']['eager_render']['value']/1e6:.2f} M;"
      f" fused kernel {rf['kernel_ms_per_step']:.3f} ms = {rf['achieved']:.0f} TFLOP/s = {rf['frac']:.3f} of burst peak {rf['peak']:.0f}")
tr = t["roofline"]
print(f"   train step {t['value']/1e3:.0f} k rays/s ({t['ms_per_step']:.2f} ms/step, e2e {t['e2e']['value']/1e3:.0f} k), {t['gpu_launches_per_step']} launches;"
      f" {tr['achieved']:.0f} TFLOP/s algorithmic = {tr['frac']:.3f} of the tensor peak; kernels {tr['kernel_ms_per_step']}")
tg = b.get("torch_gpu", {})
if "tf32" in tg:
    print(f"   unmodified reference on the same GPU: forward {tg['fp32']['forward_rays_per_s']/1e3:.1f} k (fp32) / {tg['tf32']['forward_rays_per_s']/1e3:.1f} k (TF32) rays/s,"
          f" train {tg['fp32']['train_rays_per_s']/1e3:.1f} k / {tg['tf32']['train_rays_per_s']/1e3:.1f} k; speed-ups {tg['speedup']}")
print(f"   cpu_baseline {b.get('cpu_baseline')}")
print(f"   clocks {b['clocks']}")
for n in (2, 4, 8):
    p = os.path.join(d, f"r02_bench_{n}gpu.json")
    if os.path.isfile(p):
        bn = line(p)
        print(f"== {n} GPUs: forward (weak) {bn['value']/1e6:.2f} M rays/s = {bn['value']/b['value']:.2f} x;"
              f" train weak {bn['train']['value']/1e3:.0f} k = {bn['train']['value']/t['value']:.2f} x;"
              f" train strong (global 4096) {bn['train_dp']['ms_per_step']:.2f} ms/step = {bn['train_dp']['value']/1e3:.0f} k rays/s = {t['ms_per_step']/bn['train_dp']['ms_per_step']:.2f} x,"
              f" gradient vs single GPU {bn['train_dp']['check']['grad_rel_l2_vs_single_gpu_full_batch']:.1e};"
              f" 800x800 frame {bn['frame']['ms_per_frame']:.1f} ms = {bn['frame']['value']/1e6:.1f} M rays/s, max |diff| vs single GPU {bn['frame']['max_abs_diff_vs_single_gpu_render']}")

p = os.path.join(d, "r02_launches_fused_train_step.csv")
if os.path.isfile(p):
    rows = list(csv.reader(open(p)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]; kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    names = [(r[kn].split("(")[0].replace("void ", "")[:44], float(r[mv].replace(",", ""))) for r in rows[hi + 2:]]
    start = [i for i, (n, _) in enumerate(names) if "pack_rays" in n][-1]
    agg = collections.OrderedDict()
    for n, v in names[start:]:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"== one fused training step: {len(names[start:])} launches, {tot/1e6:.2f} ms serialised (ncu, cold caches)")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:10]:
        print(f"   {n:46s} x{c:2d} {v/1e3:8.1f} us {100*v/tot:5.1f} %")

p = os.path.join(d, "r02_train_kernels_ncu_metrics.csv")
if os.path.isfile(p):
    rows = list(csv.reader(open(p)))
    h = rows[0]
    print("== training kernels, ncu --set full")
    for r in rows[2:]:
        g = dict(zip(h, r))
        print(f"   {g['Kernel Name'][:28]:28s} {float(g['gpu__time_duration.sum']):.3f} ms  tensor {float(g['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']):.1f} %"
              f"  dram R {float(g['dram__bytes_read.sum']):.2f} GB W {float(g['dram__bytes_write.sum']):.2f} GB = {float(g['dram__bytes.sum.per_second']):.2f} TB/s"
              f" ({float(g['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']):.0f} % of dram peak)  SM clock {float(g['sm__cycles_elapsed.avg.per_second']):.2f} GHz")

for name in ("parity_r02.json", "r02_parity_dropin.json"):
    p = os.path.join(d, name)
    if os.path.isfile(p):
        pj = json.load(open(p))
        print(f"== {name}")
        for k, v in pj.items():
            if isinstance(v, dict):
                flat = {a: (float(f"{x:.3g}") if isinstance(x, float) else x) for a, x in v.items() if not isinstance(x, (dict, list))}
                print(f"   {k}: {flat}")
            elif not isinstance(v, list):
                print(f"   {k}: {v}")
