"""Summarise an `ncu --set full` capture of the ADMM kernel into the text file committed under profiles/ (and refresh
profiles/admm_traffic.json, which bench.py reads for roofline.traffic).  Runs HERE (ncu -i needs no GPU).

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_admm_kernel_tm_ncu_summary.txt "header line" [--traffic]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep, out, header = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, val)}
    lines = [header, "kernel: " + m.get("Kernel Name", ("", "?"))[1]]
    for k in WANT:
        if k in m:
            lines.append("%-72s %-16s %s" % (k, m[k][0], m[k][1]))
    stalls = {h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]: float(v)
              for h, (u, v) in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
    tot = sum(stalls.values())
    lines.append("")
    lines.append("warp stall reasons (share of issue-stall ratios):")
    for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]):
        if v / tot >= 0.01:
            lines.append("  %-26s %5.1f%%" % (k, 100 * v / tot))
    # instruction mix and time share per phase: SASS instructions bucketed by how often they execute
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    isamp, iinst, isrc = sh.index("# Samples"), sh.index("Instructions Executed"), sh.index("Source")
    b = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    for r in srows[2:]:
        try:
            s, n = int(r[isamp]), int(r[iinst])
        except (ValueError, IndexError):
            continue
        t = r[isrc].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        b[n][0] += s; b[n][1] += 1; b[n][2] += n; b[n][3][op] += 1
    ts = sum(v[0] for v in b.values()); ti = sum(v[2] for v in b.values())
    lines.append("")
    lines.append("SASS instructions grouped by execution count (= phase of the solve): static instructions, share of stall samples, share of executed warp-instructions, opcode mix")
    for n, v in sorted(b.items(), key=lambda kv: -kv[1][0])[:10]:
        lines.append("  executed %9d x: %4d instr  %5.1f%% samples  %5.1f%% inst   %s" % (n, v[1], 100 * v[0] / ts, 100 * v[2] / ti,
                                                                                       " ".join("%s:%d" % kv for kv in v[3].most_common(9))))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if "--traffic" in sys.argv:
        rd = float(m["dram__bytes_read.sum"][1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m["dram__bytes_read.sum"][0]]
        wr = float(m["dram__bytes_write.sum"][1]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m["dram__bytes_write.sum"][0]]
        json.dump({"kernel": m.get("Kernel Name", ("", "?"))[1], "source": os.path.relpath(out, ROOT) + " (ncu --set full, 4096 QPs, N=30)",
                   "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr, "qps_per_launch": 4096},
                  open(os.path.join(ROOT, "profiles", "admm_traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
