"""profiles/r2_ncu.json from the full ncu capture of one 256-query step (scripts/exp_phases.py 256 refine):

    ncu -i gpurun_out/r2b_sweep.ncu-rep --page raw --csv > gpurun_out/r2b_raw.csv
    python scripts/ncu_r2_json.py gpurun_out/r2b_raw.csv gpurun_out/phases_256_r2b.log > profiles/r2_ncu.json
"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    return float(r[col[name]].replace(",", "")) if r[col[name]] not in ("", "n/a") else 0.0


events = {}
for line in open(sys.argv[2]):
    m = re.match(r"(\S+)\s+([0-9.]+) us", line)
    if m:
        events[m.group(1)] = float(m.group(2))

sweep, builder, dive, seen = [], None, None, set()
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    if name in seen:
        continue                      # the capture holds more than one step: the first of each kernel
    seen.add(name)
    d = {"kernel": name, "us": val(r, "gpu__time_duration.sum") / (1e3 if rows[1][col["gpu__time_duration.sum"]] == "ns" else 1.0),
         "requests": val(r, "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
         "sectors": val(r, "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
         "lsu_wavefronts_pct_of_peak": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
         # the pipe's peak is one wavefront per cycle per SM: wavefronts of the launch = share x cycles x SMs
         "lsu_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0 *
                           val(r, "sm__cycles_elapsed.avg") * 148,
         "l1tex_throughput_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
         "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
         "l1_hit_rate_pct": val(r, "l1tex__t_sector_hit_rate.pct"),
         "dram_read_bytes": val(r, "dram__bytes_read.sum") * (1e6 if rows[1][col["dram__bytes_read.sum"]] == "Mbyte" else 1e3 if rows[1][col["dram__bytes_read.sum"]] == "Kbyte" else 1.0),
         "dram_write_bytes": val(r, "dram__bytes_write.sum") * (1e6 if rows[1][col["dram__bytes_write.sum"]] == "Mbyte" else 1e3 if rows[1][col["dram__bytes_write.sum"]] == "Kbyte" else 1.0),
         "registers": val(r, "launch__registers_per_thread")}
    if name.startswith("k_bbg_expand"):
        sweep.append(d)
    elif name.startswith("k_bbg_dive"):
        dive = d
    elif name.startswith("k_pyramid_stream2"):
        builder = d
t = sum(d["us"] for d in sweep)
req = sum(d["requests"] for d in sweep)
sec = sum(d["sectors"] for d in sweep)
out = {
    "captured_with": "scripts/gpu_prof_r2b.sh: ncu --set full --clock-control none -k regex:k_bbg_expand|k_pyramid_stream2|"
                     "k_bbg_dive --launch-skip 104 -c 10 python scripts/exp_phases.py 256 refine (one 256-query cfg3 step "
                     "on one handle, B200 of this pool; incumbent dive after the launch of height 4)",
    "report": "gpurun_out/r2b_sweep.ncu-rep -> profiles/r2b_sweep_dive_builder_full.txt",
    "k_bbg_expand": {
        "launches": sweep, "ncu_ms": t / 1e3,
        "event_ms_at_capture": sum(v for k, v in events.items() if k.startswith("k_bbg_expand")) / 1e3,
        "requests_per_step": req, "sectors_per_step": sec, "sectors_per_request": sec / req,
        "lsu_wavefronts_per_step": sum(d["lsu_wavefronts"] for d in sweep),
        "lsu_data_pipe_frac_of_peak": sum(d["lsu_wavefronts_pct_of_peak"] * d["us"] for d in sweep) / t / 100.0,
        "lsu_data_pipe_frac_heaviest_launch": max(sweep, key=lambda d: d["us"])["lsu_wavefronts_pct_of_peak"] / 100.0,
        "issue_active_frac": sum(d["issue_active_pct"] * d["us"] for d in sweep) / t / 100.0,
        "dram_bytes_per_step": sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in sweep),
    },
    "k_bbg_dive": dive,
    "k_pyramid_stream2": None if builder is None else dict(
        builder, dram_bytes_per_launch=builder["dram_read_bytes"] + builder["dram_write_bytes"],
        event_us_at_capture=events.get("k_pyramid_stream(bounds)")),
}
print(json.dumps(out, indent=1))
