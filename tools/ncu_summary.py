"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports: one JSON record per profiled launch with the handful of
metrics DESIGN.md / profiles/README.md quote (duration, DRAM bytes, pipe utilisation, issue rate, registers, grid).

    python tools/ncu_summary.py gpurun_out/r02_prof_gemm_bwd.raw.csv [more.csv ...] > profiles/r02_ncu_summary.json

ncu prints each metric in a unit it picks per column AND the unit row is shared by the whole file, so byte / time
columns are normalised to bytes and microseconds here."""
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_bytes.sum": "l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "sm__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct": "stall_long_scoreboard_pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct": "stall_barrier_pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct": "stall_math_throttle_pct",
}


def number(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def short(name):
    name = re.sub(r"\(.*$", "", name)            # drop the argument list
    name = name.replace("magb200::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def summarise(path):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in body:
        rec = {"kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]].replace(" ", ""),
               "block": r[col["Block Size"]].replace(" ", "")}
        for metric, key in WANT.items():
            if metric in col and number(r[col[metric]]) is not None:
                rec[key] = round(number(r[col[metric]]) * SCALE.get(units[col[metric]], 1.0), 3)
        if "dram_read_bytes" in rec and "dram_write_bytes" in rec:
            rec["dram_bytes"] = rec["dram_read_bytes"] + rec["dram_write_bytes"]
            if rec.get("duration_us"):
                rec["dram_GBps"] = round(rec["dram_bytes"] / rec["duration_us"] / 1e3, 1)
        out.append(rec)
    return out


if __name__ == "__main__":
    res = {}
    for p in sys.argv[1:]:
        res[p.split("/")[-1]] = summarise(p)
    json.dump(res, sys.stdout, indent=1)
    print()
