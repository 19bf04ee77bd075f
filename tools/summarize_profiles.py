"""Print the numbers profiles/r01_summary.md quotes, from the committed raw captures (no GPU needed):
kernel shares of a step from the ncu launch list, the march kernel's tensor-pipe / clock / traffic metrics, the bench
line's roofline fields and the speed-up over the PyTorch-GPU op chain.

    python tools/summarize_profiles.py [profiles_dir] [round_prefix]
"""
import collections, csv, json, os, sys

d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
r = sys.argv[2] if len(sys.argv) > 2 else "r01"

rows = [x for x in csv.reader(open(os.path.join(d, f"{r}_launches.csv"))) if len(x) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for x in rows[1:]:
    try:
        v = float(x[vi].replace(",", ""))
    except ValueError:
        continue
    name = x[ki].split("(")[0]
    if "FillFunctor<unsigned char" in x[ki]:
        name = "[bench L2 flush] " + name
    a = agg.setdefault(name[:56], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
flush = sum(a[1] for n, a in agg.items() if n.startswith("[bench L2 flush]"))
print(f"== launch list ({sum(a[0] for a in agg.values())} launches, {tot/1e6:.2f} ms serialised; L2-flush memsets {100*flush/tot:.1f} %)")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {n:58s} x{c:3d} {v/c/1e3:8.1f} us each  {100*v/tot:5.1f} %  ({100*v/(tot-flush):5.1f} % without the flush)")

print("== march kernel, ncu --set full (coarse | fine launch)")
for x in csv.reader(open(os.path.join(d, f"{r}_march_tc_ncu_metrics.csv"))):
    if x and any(k in x[0] for k in ("gpu__time_duration", "pipe_tensor_cycles_active", "cycles_elapsed.avg.per_second", "dram__bytes_read", "lts__throughput", "registers", "shared_mem", "cluster_size")):
        print(f"  {x[0]:78s} {x[2]:>14s} | {x[3]:>14s} {x[1]}")

b = json.loads(open(os.path.join(d, f"{r}_bench.json")).read().strip().splitlines()[-1])
rf = b["roofline"]
print(f"== bench line: {b['value']/1e6:.2f} M rays/s device-resident ({b['ms_per_step']:.3f} ms/step), e2e {b['e2e']['value']/1e6:.2f} M rays/s")
print(f"   roofline: {rf['achieved']:.0f} {rf['unit']} = {rf['frac']:.3f} of the burst peak {rf['peak']:.0f}"
      + (f", {rf['frac_sustained']:.3f} of the sustained peak {rf['peak_sustained']:.0f}" if rf.get("peak_sustained") else "")
      + f"; kernel {rf['kernel_ms_per_step']:.3f} ms/step; traffic {rf['traffic']/1e6:.1f} MB/launch")
print(f"   clocks {b['clocks']}; cpu_baseline {b.get('cpu_baseline', {}).get('value')} rays/s on {b.get('cpu_baseline', {}).get('cores')} threads")
for n in (2, 4, 8):
    p = os.path.join(d, f"{r}_bench_{n}gpu.json")
    if os.path.isfile(p):
        bn = json.loads(open(p).read().strip().splitlines()[-1])
        print(f"   {n} GPUs: {bn['value']/1e6:.2f} M rays/s")
p = os.path.join(d, f"{r}_torch_gpu_reference.json")
if os.path.isfile(p):
    t = json.load(open(p))
    print(f"== PyTorch-GPU op chain on the same B200: fp32 {t['fp32']['rays_per_s']/1e3:.1f} k rays/s, tf32 {t['tf32']['rays_per_s']/1e3:.1f} k rays/s"
          f" -> fused / tf32 = {b['value']/t['tf32']['rays_per_s']:.1f} x, fused / fp32 = {b['value']/t['fp32']['rays_per_s']:.1f} x")
