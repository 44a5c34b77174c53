#!/usr/bin/env python
"""Turn one GPU round's ncu outputs (gpurun_out/) into the committed summary under profiles/.

  python profiles/summarize.py <tag> [<out-prefix>]

Reads   gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum: every launch, cold cache, serialised)
        gpurun_out/traffic64_<tag>.csv  (duration + dram bytes per launch at the bench batch size E=64)
        gpurun_out/prof_<tag>.ncu-rep   (ncu --set full of the write and read kernels, E=16)
        gpurun_out/bench_<tag>.log      (the bench line of the same build, taken WITHOUT a profiler)
Writes  profiles/<prefix>_launches.csv, profiles/<prefix>_traffic64.csv (copies), profiles/<prefix>_summary.md,
        profiles/write_kernel_traffic.json (the `roofline.traffic` figure bench.py reports).
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KERNELS = ("backproject", "read_pool", "frame_count", "expand", "write_mean", "finalize", "reset_touched")


def load_metrics(fn):
    rows = [r for r in csv.reader(l for l in open(fn) if l.startswith('"'))]
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        d.setdefault((int(r[ii]), name), {})[r[mi]] = float(r[vi].replace(",", ""))
    return d


def rep_section(rep, title, md, extra=()):
    """Key metrics + stall breakdown of every launch captured in one .ncu-rep (ncu --set full)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size",
            "launch__grid_size", "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"] + list(extra)
    md += [f"## {title}", ""]
    ki = h.index("Kernel Name")
    md += ["| metric | " + " | ".join(f"`{r[ki].split('(')[0].replace('<unnamed>::', '').replace('void ', '')[:28]}`" for r in rows[2:]) + " |",
           "|---|" + "---|" * len(rows[2:])]
    for w in want:
        if w in h:
            i = h.index(w)
            md.append(f"| {w} [{units[i]}] | " + " | ".join(r[i][:14] for r in rows[2:]) + " |")
    md.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks, names, cur = [], [], None
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append(cur)
            names.append(r[1].split("(")[0].replace("<unnamed>::", "").replace("void ", ""))
            continue
        if cur is not None:
            cur.append(r)
    md += ["Warp-stall sampling (source page, all samples), first capture of each kernel:", ""]
    done = set()
    for name, b in zip(names, blocks):
        if name in done or len(b) < 2:
            continue
        done.add(name)
        hh, data = b[0], b[1:]
        si = hh.index("# Samples")
        stalls = [c for c in hh if c.startswith("stall_") and "Not Issued" not in c]
        tot = {s: sum(int(r[hh.index(s)]) for r in data) for s in stalls}
        total = sum(int(r[si]) for r in data) or 1
        md.append(f"- `{name}`: " + ", ".join(f"{s[6:]} {v / total:.2f}" for s, v in sorted(tot.items(), key=lambda x: -x[1])[:7])
                  + f" ({len(data)} SASS instructions)")
    md.append("")


def main():
    tag = sys.argv[1]
    prefix = sys.argv[2] if len(sys.argv) > 2 else tag
    md = [f"# ncu summary `{prefix}` (gpurun tag `{tag}`)", "",
          "All numbers below were taken under `ncu --clock-control none` (cold cache, serialised launches): compare SHARES and "
          "bytes, not absolute times.  The timing authority is `bench.py` (CUDA events, no profiler); its line for the same "
          "build is quoted at the end.", ""]

    fn = os.path.join(OUT, f"launches_{tag}.csv")
    if os.path.exists(fn):
        shutil.copy(fn, os.path.join(ROOT, "profiles", f"{prefix}_launches.csv"))
        agg = collections.OrderedDict()
        for (_, k), v in load_metrics(fn).items():
            if any(x in k for x in KERNELS):
                agg.setdefault(k, []).append(v["gpu__time_duration.sum"])
        tot = sum(sum(v) for v in agg.values())
        md += ["## Launch list (`profiles/prof_step.py`, E=16 episodes x 3 frame-steps, 480x640, C=256, 500x500 grid)", "",
               "| kernel | launches | mean us | share of the frame-step |", "|---|---|---|---|"]
        for k, v in agg.items():
            md.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |")
        md.append("")

    fn = os.path.join(OUT, f"traffic64_{tag}.csv")
    traffic = None
    if os.path.exists(fn):
        shutil.copy(fn, os.path.join(ROOT, "profiles", f"{prefix}_traffic64.csv"))
        md += ["## DRAM traffic per launch at the bench batch size (E=64 episodes, one frame-step per launch)", "",
               "| kernel | us | dram read MB | dram write MB |", "|---|---|---|---|"]
        seen = collections.OrderedDict()
        for (_, k), v in load_metrics(fn).items():
            if any(x in k for x in KERNELS) and "dram__bytes_read.sum" in v:
                seen.setdefault(k, []).append(v)
        for k, vs in seen.items():
            v = vs[-1]                                   # last launch: grids populated
            md.append(f"| `{k}` | {v['gpu__time_duration.sum'] / 1e3:.1f} | {v['dram__bytes_read.sum'] / 1e6:.1f} | {v['dram__bytes_write.sum'] / 1e6:.1f} |")
            if "write_mean" in k:
                traffic = {"kernel": k, "dram_bytes_read": v["dram__bytes_read.sum"], "dram_bytes_write": v["dram__bytes_write.sum"],
                           "traffic_bytes_per_launch": v["dram__bytes_read.sum"] + v["dram__bytes_write.sum"],
                           "episodes_per_launch": 64, "source": f"profiles/{prefix}_traffic64.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum)"}
        md.append("")
        if traffic:
            E, N, C = 64, 480 * 640, 256
            alg = E * (N * C * 4 + N * 4)
            md += [f"Write kernel: DRAM traffic {traffic['traffic_bytes_per_launch'] / 1e9:.3f} GB per launch vs algorithmic "
                   f"{alg / 1e9:.3f} GB (features + index plane) = x{traffic['traffic_bytes_per_launch'] / alg:.3f}.", ""]
            json.dump(traffic, open(os.path.join(ROOT, "profiles", "write_kernel_traffic.json"), "w"), indent=1)

    rep = os.path.join(OUT, f"prof_{tag}.ncu-rep")
    if os.path.exists(rep):
        rep_section(rep, "`ncu --set full` of the write and read kernels (E=16), key metrics per captured launch", md)

    fn = os.path.join(OUT, f"prof_fuse_{tag}.json")
    if os.path.exists(fn):
        md += ["## Projection + fusion stage (`profiles/prof_fuse.py`, E=64, three levels, K=512 -> N=256; CUDA events, no profiler)", "",
               "```", open(fn).read().strip().splitlines()[-1], "```", ""]
    rep = os.path.join(OUT, f"prof_fuse_{tag}.ncu-rep")
    if os.path.exists(rep):
        rep_section(rep, "`ncu --set full` of `project_fuse_persistent_kernel` (E=64, all levels in one launch)", md,
                    extra=["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__cycles_elapsed.max"])

    fn = os.path.join(OUT, f"configs_{tag}.log")
    if os.path.exists(fn):
        md += ["## Other BASELINE configurations (`profiles/bench_configs.py`; CUDA events, no profiler)", "", "```"] + \
              [l for l in open(fn).read().strip().splitlines() if l.startswith("{")] + ["```", ""]
    fn = os.path.join(OUT, f"layouts_{tag}.json")
    if os.path.exists(fn):
        md += ["## Dense write by feature layout (`profiles/prof_layouts.py`, E=64, C=256; CUDA events, no profiler)", "",
               "```", open(fn).read().strip().splitlines()[-1], "```", ""]
    fn = os.path.join(OUT, f"objects_{tag}.json")
    if os.path.exists(fn):
        md += ["## Object regime, stage breakdown (`profiles/prof_objects.py`, E=64, C=512, <=16 detections per frame; ms per launch)", "",
               "```", open(fn).read().strip().splitlines()[-1], "```", ""]
    fn = os.path.join(OUT, f"launches_obj_{tag}.csv")
    if os.path.exists(fn):
        shutil.copy(fn, os.path.join(ROOT, "profiles", f"{prefix}_launches_objects.csv"))
        agg = collections.OrderedDict()
        for (_, k), v in load_metrics(fn).items():
            agg.setdefault(k, []).append(v["gpu__time_duration.sum"])
        tot = sum(sum(v) for v in agg.values())
        md += ["Launch list of the same script under ncu (E=16, cold cache, serialised):", "", "| kernel | launches | mean us | share |", "|---|---|---|---|"]
        for k, v in agg.items():
            md.append(f"| `{k[:60]}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / tot:.3f} |")
        md.append("")

    stress = [(k, os.path.join(OUT, f"stress_{k}_{tag}.log")) for k in ("read", "paste", "geometry", "dense_write", "objects", "fuse")]
    if any(os.path.exists(f) for _, f in stress):
        md += ["## Adversarial / randomised sweeps against the oracle (`profiles/stress_*.py`)", ""]
        for k, f in stress:
            if os.path.exists(f):
                md.append(f"- `stress_{k}.py`: " + open(f).read().strip().splitlines()[-1])
        md.append("")

    fn = os.path.join(OUT, f"bench_{tag}.log")
    if os.path.exists(fn):
        line = open(fn).read().strip().splitlines()[-1]
        md += ["## bench.py line of the same build (no profiler)", "", "```", line, "```", ""]
    open(os.path.join(ROOT, "profiles", f"{prefix}_summary.md"), "w").write("\n".join(md))
    print("\n".join(md[:60]))


if __name__ == "__main__":
    main()
