"""Turns the round's ncu outputs under gpurun_out/ into the committed summaries under profiles/:
   r02_full_<w>_raw.csv  -> profiles/r02_ncu_full_<kernel>_<w>.md   (selected metrics of the --set full capture)
   r02_launches_<w>.csv  -> profiles/r02_launches_<w>.md            (per-kernel share of the device time of one run)"""
import collections
import csv
import io
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]
WHAT = {"c2": ("knn_filter", "BASELINE config 2 at the bench's own launch size: 1M x 16 f32 points, 1 000 000 queries, k = 10 (`scripts/ncu_targets.py c2`; seeded dense plan: main launch of 13 whole waves = 985 088 queries, then the tail launch)"),
        "t128": ("knn_filter", "north-star shape 10M x 128 f32, 75 776 queries = two whole waves of 148 CTAs x 256 queries, k = 10 (`scripts/ncu_targets.py t128`)"),
        "c3": ("knn_filter", "VantagePointTree 1M x 64 f32 mixture, 303 104 queries, query_nearest: the seeded tensor scan on the ball partition (`scripts/ncu_targets.py c3`)"),
        "c3p": ("pruned_scan", "VantagePointTree 1M x 64 f32 mixture, 303 104 queries, query_nearest: the PRUNED tensor scan on the two-means partition -- seeds, tile bitmaps (ball_tile_kernel: the balls of all warps, then the queries of the wide warps), filter over the tiles the bitmaps leave (`scripts/ncu_targets.py c3`)"),
        "c4": ("radius", "BallTree::query_radius 10M x 3 f32, r = 0.01, 262 144 queries in one chunk (`scripts/ncu_targets.py c4`)"),
        "c4knn": ("knn_warp", "BallTree::query 10M x 3 f32, 1 000 000 device-resident queries, k = 10: the warp-per-query scan (`scripts/ncu_targets.py c4knn`)"),
        "c1": ("knn_warp", "BASELINE config 1: BallTree 10k x 3 f64, every point a query (query_self), k = 10 (`scripts/ncu_targets.py c1`)")}
for w, (kern, desc) in WHAT.items():
    path = os.path.join(G, f"r02_full_{w}_raw.csv")
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = [f"# Round 2 -- `ncu --set full --clock-control none` of `{kern}` kernels: {desc}", ""]
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        if d.get("gpu__time_duration.sum", "").startswith("-nan") or d.get("dram__bytes_read.sum", "nan").endswith("nan"):
            out += [f"(launch `{d.get('Kernel Name', '')[:80]}`, {d.get('gpu__time_duration.sum', '?')} {u.get('gpu__time_duration.sum', '')}: metric passes incomplete, omitted)", ""]
            continue
        out += [f"`{d.get('Kernel Name', '')}`", "", "| metric | unit | value |", "|---|---|---|"]
        for k in KEYS:
            if k in d:
                out.append(f"| {k} | {u[k]} | {d[k]} |")
        out.append("")
    open(os.path.join(P, f"r02_ncu_full_{kern}_{w}.md"), "w").write("\n".join(out) + "\n")
    print("wrote", f"r02_ncu_full_{kern}_{w}.md")
for w in ("c2", "c3", "c3p", "c4", "c4knn", "c1"):
    path = os.path.join(G, f"r02_launches_{w}.csv")
    if not os.path.exists(path):
        continue
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("".join(lines))))
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for x in rows[1:]:
        if len(x) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", x[ci["Kernel Name"]]).replace("void ", "")
        v = float(x[ci["Metric Value"]].replace(",", ""))
        unit = x[ci["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += v
    mine = {k: v for k, v in agg.items() if not k.startswith("at::")}
    tot = sum(v for _, v in mine.values())
    out = [f"# Round 2 -- launch list of `python scripts/ncu_targets.py {w.replace('c3p', 'c3')}` (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
           f"{WHAT[w][1]}.  Two calls plus the tree build; per-launch times are cold-cache and serialised, so only the SHARES matter.",
           f"Library kernels only ({len(mine)} kernels, {sum(c for c, _ in mine.values())} launches, {tot:.2f} ms); "
           f"torch's generator kernels of the synthetic inputs ({sum(c for k, (c, _) in agg.items() if k.startswith('at::'))} launches) are left out.", "",
           "| kernel | launches | ms | share |", "|---|---|---|---|"]
    for k, (c, v) in sorted(mine.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:100]}` | {c} | {v:.3f} | {100 * v / tot:.2f} % |")
    open(os.path.join(P, f"r02_launches_{w}.md"), "w").write("\n".join(out) + "\n")
    print("wrote", f"r02_launches_{w}.md")
