"""Condense ncu output into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches_summary.txt
  python tools/ncu_summary.py full gpurun_out/prof_pixgemm_r1.ncu-rep > profiles/r1_pixgemm_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["Grid Size", "Block Size", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1.0)
        agg[r[ki][:110]][0] += 1
        agg[r[ki][:110]][1] += v
        tot += v
    print("ncu --metrics gpu__time_duration.sum --clock-control none: %d launches, %.3f ms total (cold-cache, serialised)" % (len(data), tot))
    print("%10s %7s %6s  %s" % ("total_ms", "share", "count", "kernel"))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.3f %6.2f%% %6d  %s" % (v, 100 * v / tot, n, k))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("ncu --set full --clock-control none: %s (%d launches)" % (path, len(data)))
    for r in data:
        print("--- %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("    %-66s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
