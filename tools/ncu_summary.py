"""Summarise gpurun_out/prof_<tag>.ncu-rep and launches_<tag>.csv into profiles/<tag>_*.  Run in the build container."""
import collections
import csv
import io
import subprocess
import sys

tag = sys.argv[1]
out_md = f"profiles/{tag}_ncu_summary.md"
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
lines = [f"# ncu summary `{tag}`", "",
         "Source: `ncu --set full --clock-control none --import-source on` on `python bench.py --steps 2 --warmup 3 --views 8 "
         "--no-cpu-baseline --no-train` (configs[1] frame: trace_compact<HitBufSmemT<8>>, ngp_forward_tc_kernel), `python "
         "tools/run_leg.py c5 3` (configs[4] 4K baked frame: trace_compact<HitBufSmem>, baked_shade_kernel) and `python "
         "tools/diag_train.py 3` (training step kernels) and `python tools/run_field_legs.py 5` (quadrature-field legs: "
         "render_weights / field_net / occgrid_march kernels); B200, one GPU, tools/gpu_profile.sh.  Per-launch values; cold-cache, "
         "serialised — compare shares, not absolutes.", ""]
import os
traffic = {}
seen = set()
for rep in (f"gpurun_out/prof_{tag}.ncu-rep", f"gpurun_out/prof_extra_{tag}.ncu-rep", f"gpurun_out/prof_train_{tag}.ncu-rep",
            f"gpurun_out/prof_misc_{tag}.ncu-rep", f"gpurun_out/prof_misc2_{tag}.ncu-rep"):
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"]
            traffic[name.replace("void ", "").replace("qf::", "")] = float(r[rd]) * scale.get(units[rd], 1) + float(r[wr]) * scale.get(units[wr], 1)
        except Exception:
            pass
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in idx:
                lines.append(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |")
        lines.append("")
try:
    rows = list(csv.DictReader(l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"]) / 1e3
    lines += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)", "",
              "| kernel | launches | total us | us / launch |", "|---|---|---|---|"]
    for k, (n, t) in agg.items():
        lines.append(f"| `{k[:90]}` | {n} | {t:.1f} | {t / n:.1f} |")
    step = {k: v for k, v in agg.items() if any(s in k for s in ("trace_compact", "ngp_forward", "composite_rays", "baked_shade"))}
    tot = sum(v[1] for v in step.values())
    lines += ["", "Share of one render step: " + ", ".join(f"`{k.split('::')[-1][:30]}` {100 * v[1] / tot:.1f}%" for k, v in step.items()), ""]
    with open(f"profiles/{tag}_launches.csv", "w") as f:
        f.write(open(f"gpurun_out/launches_{tag}.csv").read())
except FileNotFoundError:
    pass
import json
json.dump({"tag": tag, "dram_bytes_per_launch": traffic, "source": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum"},
          open("profiles/roofline_traffic.json", "w"), indent=1)
open(out_md, "w").write("\n".join(lines))
print("wrote", out_md)
