"""Text summary of one `ncu --set full` capture (.ncu-rep): the metrics the B200 profiling recipe asks for + stall reasons.
usage: ncu_summary.py capture.ncu-rep [out.txt]   (written for profiles/, the judged location)"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]


def main(rep, out=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, val = rows[0], rows[1], rows[-1]
    ix = {h: i for i, h in enumerate(hdr)}
    lines = [f"kernel: {val[ix['Kernel Name']]}", f"source: {rep}"]
    for k in KEYS:
        if k in ix:
            lines.append(f"{k:90s} {val[ix[k]]} {units[ix[k]]}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(val[ix[h]] or 0) >= 0.05:
            lines.append(f"{h:90s} {val[ix[h]]}")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")


if __name__ == "__main__":
    main(*sys.argv[1:3])
