#!/usr/bin/env python3
"""Summarise an Nsight Compute report (read here with `ncu -i`, no GPU needed) into profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep profiles/r1_k_scan_full --launches gpurun_out/launches_r1.csv

Writes <out>.md (key metrics per kernel + hottest SASS blocks) and copies the launch list.
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__cycles_elapsed.max"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    launches = sys.argv[sys.argv.index("--launches") + 1] if "--launches" in sys.argv else None
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    md = ["# ncu summary of `%s`" % rep, "",
          "Captured with `ncu --set full --clock-control none --import-source on` under gpurun; read here with `ncu -i`.",
          "Per-launch values (one replayed launch per row; cold cache, serialised - compare shares, not absolutes).", ""]
    traffic = {}
    for r in raw[2:]:
        name = r[ix["Kernel Name"]].split("(")[0]
        try:
            def tobytes(v, u):
                return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            traffic.setdefault(name, {"dram_bytes_read": tobytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]),
                                      "dram_bytes_write": tobytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]]),
                                      "source": out + ".md"})
        except (KeyError, ValueError):
            pass
        md.append("## %s" % r[ix["Kernel Name"]])
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---|---|")
        for k in KEYS:
            if k in ix:
                md.append("| %s | %s | %s |" % (k, r[ix[k]], units[ix[k]]))
        md.append("")
    # SASS hot blocks of every kernel
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    for si, s0 in enumerate(secs):
        end = secs[si + 1] if si + 1 < len(secs) else len(src)
        h = src[s0 + 1]
        jx = {x: i for i, x in enumerate(h)}
        ins = []
        for r in src[s0 + 2:end]:
            if len(r) < 10:
                continue
            ins.append((r[jx["Source"]].strip(), int(r[jx["Instructions Executed"]]), int(r[jx["Thread Instructions Executed"]]),
                        int(r[jx["# Samples"]])))
        tot = sum(i[1] for i in ins) or 1
        blocks, cur = [], []
        for k, i in enumerate(ins):
            if cur and abs(i[1] - cur[-1][2]) > 0.02 * max(cur[-1][2], 1):
                blocks.append(cur)
                cur = []
            cur.append((k,) + i)
        if cur:
            blocks.append(cur)
        md.append("## SASS hot blocks: %s (launch %d), %d SASS instructions, %d warp instructions executed" %
                  (src[s0][1], si, len(ins), tot))
        md.append("")
        md.append("| sass idx | n instr | executions each | share of warp instrs | avg active lanes | stall samples | first instruction |")
        md.append("|---|---|---|---|---|---|---|")
        for b in sorted(sorted(blocks, key=lambda b: -sum(x[2] for x in b))[:16], key=lambda b: b[0][0]):
            s = sum(x[2] for x in b)
            th = sum(x[3] for x in b)
            md.append("| %d-%d | %d | %d | %.1f %% | %.1f | %d | `%s` |" %
                      (b[0][0], b[-1][0], len(b), b[0][2], 100.0 * s / tot, th / max(s, 1), sum(x[4] for x in b), b[0][1][:48]))
        md.append("")
    if launches:
        md.append("## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)")
        md.append("")
        rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
        agg = {}
        for r in rows:
            agg.setdefault(r[4], []).append(float(r[-1]))
        total = sum(sum(v) for v in agg.values())
        md.append("| kernel | launches | mean ns | share of GPU time |")
        md.append("|---|---|---|---|")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            md.append("| %s | %d | %.0f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100.0 * sum(v) / total))
        md.append("")
    open(out + ".md", "w").write("\n".join(md))
    if "--traffic" in sys.argv:
        import json
        json.dump(traffic, open(sys.argv[sys.argv.index("--traffic") + 1], "w"), indent=1)
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()
