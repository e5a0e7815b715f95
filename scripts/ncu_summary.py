"""Summarise ncu outputs into small text files for profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv > profiles/r1_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof_bb.ncu-rep > profiles/r1_k_bb_score.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum",
    "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        a = agg.setdefault((name, row["Grid Size"], row["Block Size"]), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): shares, not absolutes")
    print("%-28s %-18s %-14s %6s %12s %10s %7s" % ("kernel", "grid", "block", "n", "total_us", "avg_us", "share"))
    for (k, g, b), a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-28s %-18s %-14s %6d %12.1f %10.2f %7.3f" % (k, g, b, a[0], a[1], a[1] / a[0], a[1] / tot))
    by_name = collections.OrderedDict()
    for (k, g, b), a in agg.items():
        c = by_name.setdefault(k, [0, 0.0])
        c[0] += a[0]
        c[1] += a[1]
    print("\n# per kernel name")
    for k, a in sorted(by_name.items(), key=lambda x: -x[1][1]):
        print("%-28s n=%5d total=%10.1f us share=%.3f" % (k, a[0], a[1], a[1] / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full --clock-control none, one column per captured launch (%s)" % path)
    print("%-72s %-10s %s" % ("kernel", "", "  ".join(r[idx["Kernel Name"]].split("(")[0] for r in rows[2:])))
    for k in KEYS:
        if k in idx:
            print("%-72s %-10s %s" % (k, units[idx[k]], "  ".join(r[idx[k]] for r in rows[2:])))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
