"""Summarise an .ncu-rep (ncu --set full) into a small text file for profiles/ (run where ncu is installed).

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.txt ["note"]
"""
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = ["# %s" % rep, "# %s" % note, ""]
    for r in rows[2:]:
        lines.append("kernel: " + re.sub(r"\(.*", "", r[hdr.index("Kernel Name")]))
        for k in KEYS:
            if k in hdr:
                lines.append("  %-68s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(r[i]) for i, h in enumerate(hdr)
                  if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and r[i].isdigit()}
        tot = sum(stalls.values()) or 1
        lines.append("  warp-state samples: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot)
                                                           for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
        lines.append("")
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(sass)))
    if len(srows) > 2 and "Source" in srows[1]:
        h = srows[1]
        i_src, i_ex = h.index("Source"), h.index("Instructions Executed")
        ops = {}
        for r in srows[2:]:
            if len(r) < len(h) or not r[i_ex].isdigit():
                break
            op = re.sub(r"^@!?U?P\d+\s+", "", r[i_src].strip()).split()[0].split(".")[0]
            ops[op] = ops.get(op, 0) + int(r[i_ex])
        tot = sum(ops.values()) or 1
        lines.append("SASS opcode mix (first kernel, executed warp instructions): " +
                     ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
        blackwell = [k for k in ops if k.startswith(("UTC", "LDTM", "STTM", "UBLKCP", "UTMA", "LDGSTS"))]
        lines.append("Blackwell / async mnemonics present: " + ", ".join(sorted(blackwell)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
