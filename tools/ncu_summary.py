"""CPU-side helper: the handful of numbers DESIGN.md and bench.py quote from an `ncu --set full` capture.

    python tools/ncu_summary.py gpurun_out/prof_frames.ncu-rep > profiles/frames_ncu_summary.json
Reads the report's raw page (`ncu -i X.ncu-rep --page raw --csv`) and prints one JSON object for its first kernel.
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "gpu_time_duration",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_rate_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_bytes",
    "launch__shared_mem_per_block_static": "static_smem_bytes",
    "launch__occupancy_limit_registers": "occupancy_limit_registers",
    "sm__maximum_warps_per_active_cycle_pct": "theoretical_occupancy_pct",
}


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    header, units, data = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(header)}
    res = {"report": rep, "kernel": data[col["Kernel Name"]], "grid": data[col["Grid Size"]], "block": data[col["Block Size"]]}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte/block": 1.0, "Kbyte/block": 1e3,
             "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    for metric, name in WANT.items():
        if metric in col:
            v = data[col[metric]].replace(",", "")
            unit = units[col[metric]]
            try:
                res[name] = float(v) * scale.get(unit, 1.0)  # bytes and seconds, whatever unit ncu chose to print
            except ValueError:
                res[name] = v
    if "gpu_time_duration" in res:
        res["gpu_time_duration_s"] = res.pop("gpu_time_duration")
    for k, v in (kv.split("=", 1) for kv in sys.argv[2:]):  # free-form annotations: source=..., workload=..., note=...
        res[k] = v
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
