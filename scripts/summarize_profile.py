#!/usr/bin/env python
"""Summarise an ncu launch list (csv) and an `ncu --set full` report of the solve kernel into profiles/<tag>_*.{md,json}.

  python scripts/summarize_profile.py <tag> [gpurun_out/launches_<tag>.csv] [gpurun_out/prof_<tag>.ncu-rep]

Run on the CPU box (ncu -i reads the report without a GPU)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % tag)
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
# optional: what the full capture is of (default: the faithful solve kernel on the bench workload, which also refreshes profiles/traffic.json)
what = sys.argv[4] if len(sys.argv) > 4 else "one `cbfnmpc_lane_kernel` launch (4096 cfg-3 instances, `scripts/dev/bench_quick.py 4096`)"
is_headline = len(sys.argv) <= 4
out_md = [("# ncu summary `%s`\n" % tag)]
summary = {"tag": tag}

if os.path.exists(launches):
    rows = [r for r in csv.reader(l for l in open(launches) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = r[ik].split("(")[0][:90]
        agg[name][0] += 1; agg[name][1] += float(r[iv].replace(",", "")) * 1e-6
    tot = sum(v[1] for v in agg.values())
    out_md.append("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `python bench.py --steps 3 --warmup 3 --no-cpu-baseline`)\n")
    out_md.append("cold-cache, serialised per launch: compare SHARES, not absolutes.\n")
    out_md.append("| kernel | launches | total ms | share | mean ms |\n|---|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out_md.append("| `%s` | %d | %.3f | %.1f %% | %.4f |" % (k, n, t, 100 * t / tot, t / n))
    summary["launches"] = {k: {"n": n, "ms": t, "share": t / tot} for k, (n, t) in agg.items()}
    out_md.append("")

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    h, u, v = rr[0], rr[1], rr[2]
    m = {a: (c, b) for a, b, c in zip(h, u, v)}
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
            "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
            "sm__icc_request_hit_rate.pct", "sm__cycles_active.avg"]
    out_md.append("## `ncu --set full --clock-control none` of %s\n" % what)
    out_md.append("| metric | value | unit |\n|---|---|---|")
    summary["metrics"] = {}
    for k in keys:
        if k in m:
            out_md.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
            summary["metrics"][k] = m[k][0]
    stall = {a.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(c.replace(",", ""))
             for a, c in zip(h, v) if a.startswith("smsp__average_warps_issue_stalled_") and a.endswith("_per_issue_active.ratio")}
    tot = sum(stall.values())
    out_md.append("\n### Warp states per issue-active cycle (sum %.2f warps)\n" % tot)
    out_md.append("| state | warps | share |\n|---|---|---|")
    for k, x in sorted(stall.items(), key=lambda kv: -kv[1]):
        if x > 0.005: out_md.append("| %s | %.3f | %.1f %% |" % (k, x, 100 * x / tot))
    summary["stall"] = stall
    try:
        rd = float(m["dram__bytes_read.sum"][0].replace(",", "")); wr = float(m["dram__bytes_write.sum"][0].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        summary["traffic_bytes"] = rd * scale[m["dram__bytes_read.sum"][1]] + wr * scale[m["dram__bytes_write.sum"][1]]
        summary["traffic_read_bytes"] = rd * scale[m["dram__bytes_read.sum"][1]]
        summary["traffic_write_bytes"] = wr * scale[m["dram__bytes_write.sum"][1]]
    except Exception as e:
        print("traffic:", e)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sr = list(csv.reader(io.StringIO(src)))
    hdr = sr[1]
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    c, s = Counter(), Counter(); tot = 0; tots = 0
    for r in sr[2:]:
        if len(r) < 10: continue
        op = r[isrc].split(); o = (op[0] if not op[0].startswith("@") else op[1]).split(".")[0]
        c[o] += int(r[ia]); s[o] += int(r[isamp]); tot += int(r[ia]); tots += int(r[isamp])
    out_md.append("\n### SASS opcode mix, weighted by executions (%d warp instructions, %d static)\n" % (tot, len(sr) - 2))
    out_md.append("| opcode | executed | share | stall samples share |\n|---|---|---|---|")
    for k, x in c.most_common(24):
        out_md.append("| %s | %d | %.1f %% | %.1f %% |" % (k, x, 100 * x / tot, 100 * s[k] / max(tots, 1)))
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    out_md.append("\nFP64 arithmetic (DFMA+DMUL+DADD) = %.1f %% of executed warp instructions; local-memory traffic (LDL+STL) = %.1f %%.\n"
                  % (100 * fp / tot, 100 * (c["LDL"] + c["STL"]) / tot))
    summary["inst_total"] = tot; summary["inst_static"] = len(sr) - 2; summary["fp64_share"] = fp / tot; summary["local_share"] = (c["LDL"] + c["STL"]) / tot

os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
if "traffic_bytes" in summary and is_headline:
    json.dump({"traffic_bytes": summary["traffic_bytes"], "read_bytes": summary["traffic_read_bytes"], "write_bytes": summary["traffic_write_bytes"],
               "kernel_version": tag,
               "source": "profiles/%s_ncu_summary.md: dram__bytes_read.sum + dram__bytes_write.sum of one cbfnmpc_lane_kernel launch (4096 cfg-3 instances, ncu --set full)" % tag},
              open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % tag), "w").write("\n".join(out_md) + "\n")
json.dump(summary, open(os.path.join(ROOT, "profiles", "%s_ncu_summary.json" % tag), "w"), indent=1)
print("\n".join(out_md))
