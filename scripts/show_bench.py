"""Print the essentials of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
for k in ("impl", "value", "ms_per_step", "n_gpus", "host_issue_ms_per_step", "gpu_launches", "repeats", "per_rank",
          "warm", "check", "clocks", "cpu_baseline", "strong"):
    if k in d: print(k, d[k])
print("e2e", d["e2e"].get("value"), d["e2e"].get("ms_per_step"))
if "roofline" in d:
    r = d["roofline"]
    print("roofline", {k: r[k] for k in ("bound", "achieved", "frac", "ms_per_step", "share_of_step", "launches",
                                          "children_scored_per_step", "groups_per_list") if k in r})
    print("hbm_equivalent", r.get("hbm_equivalent"))
    print("pyr", d["roofline_pyramid"]["ms_per_step"], d["roofline_pyramid"]["frac"])
    print("phases", d["phases"]["kernel_ms"])
    print("e2e_cpu_final_matcher_ms_per_step", d.get("e2e_cpu_final_matcher_ms_per_step"))
if "single_scan" in d:
    for k, v in d["single_scan"].items():
        print("single", k, v if not isinstance(v, dict) else {a: b for a, b in v.items() if a in ("gpu_e2e", "cpu_1core", "cpu_1core_scaled", "ratio")})
