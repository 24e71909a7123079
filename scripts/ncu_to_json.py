"""profiles/*.json that bench.py reads, from `ncu -i X.ncu-rep --page raw --csv` files (one row per captured launch).

    python scripts/ncu_to_json.py orb  gpurun_out/r02t_orb_level_kernel_raw.csv  profiles/orb_level_traffic.json
    python scripts/ncu_to_json.py lk   gpurun_out/r02t_lk_track2_kernel_raw.csv  profiles/lk_track_issue.json
"""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "inst": 1.0, "": 1.0,
        "inst/cycle": 1.0, "%": 1.0}


def load(path):
    r = list(csv.reader(open(path)))
    hdr, units, rows = r[0], r[1], r[2:]
    idx = {n: i for i, n in enumerate(hdr)}

    def get(row, name):
        v = row[idx[name]].replace(",", "")
        try:
            return float(v) * UNIT.get(units[idx[name]], 1.0)
        except ValueError:
            return v
    return rows, get


kind, src, dst = sys.argv[1:4]
rows, get = load(src)
if kind == "orb":
    per = [{"grid": get(r, "Grid Size"), "duration_us": get(r, "gpu__time_duration.sum"),
            "dram_read": get(r, "dram__bytes_read.sum"), "dram_write": get(r, "dram__bytes_write.sum"),
            "warp_instructions": get(r, "smsp__inst_executed.sum"),
            "ipc_active": get(r, "sm__inst_executed.avg.per_cycle_active")} for r in rows]
    out = {"dram_bytes_per_launch_group": sum(p["dram_read"] + p["dram_write"] for p in per),
           "launch_group": "the 8 orb_level_kernel launches of one step (32 frames of 1241x376)", "frames_per_launch_group": 32,
           "source": f"ncu --set full --clock-control none, {src.split('/')[-1]} (dram__bytes_read.sum + dram__bytes_write.sum, "
                     "smsp__inst_executed.sum per launch)",
           "note": "below the algorithmic 215 MB because the 126 MB L2 keeps part of the pyramid between the level launches "
                   "and holds dirty lines past the kernel end",
           "warp_instructions_per_launch_group": sum(p["warp_instructions"] for p in per),
           "duration_us_under_ncu": sum(p["duration_us"] for p in per), "per_launch": per}
elif kind == "lk":
    r = rows[0]
    out = {"kernel": "lk_track2_kernel<PERSIST>", "streams": 32, "points_per_launch": 64000,
           "warp_instructions_per_launch": get(r, "smsp__inst_executed.sum"),
           "duration_us_under_ncu": get(r, "gpu__time_duration.sum"),
           "issue_slots_busy_pct_under_ncu": get(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") if False else None,
           "pipe_utilisation_pct": {"alu": get(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                                    "fma_imad_idp": get(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                                    "lsu": get(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                                    "xu": get(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                                    "adu": get(r, "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active")},
           "issue_active_pct": get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
           "dram_bytes": get(r, "dram__bytes_read.sum") + get(r, "dram__bytes_write.sum"),
           "source": f"ncu --set full --clock-control none of the bench workload (32 streams x 2000 points, 1241x376), "
                     f"{src.split('/')[-1]}: smsp__inst_executed.sum"}
    out.pop("issue_slots_busy_pct_under_ncu")
else:
    raise SystemExit("kind: orb | lk")
json.dump(out, open(dst, "w"), indent=1)
print(dst, {k: v for k, v in out.items() if not isinstance(v, (list, dict))})
