#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full capture) into the small JSON files kept under profiles/.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json "note ..."

Runs where ncu is installed (no GPU needed): `ncu -i <rep> --page raw --csv`."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "lts__t_sectors.avg.pct_of_peak_sustained_elapsed": "l2_sectors_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__cycles_active.avg": "sm_cycles_active",
    "sm__cycles_elapsed.avg.per_second": "sm_clock_hz",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
              "usecond": 1.0, "nsecond": 1e-3}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kernels = []
    for r in data:
        k = {"name": r[hdr.index("Kernel Name")][:90]}
        for h, u, v in zip(hdr, units, r):
            if h in KEEP:
                try:
                    val = float(v.replace(",", ""))
                except ValueError:
                    continue
                if KEEP[h].startswith("dram_") and KEEP[h] != "dram_throughput_pct":
                    val *= UNIT_SCALE.get(u, 1.0)
                    k[KEEP[h] + "_bytes"] = val
                elif KEEP[h] == "duration_us":
                    k["duration_us"] = val * UNIT_SCALE.get(u, 1.0)
                else:
                    k[KEEP[h]] = val
            elif h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    k.setdefault("stall_per_issue", {})[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(float(v), 3)
                except ValueError:
                    pass
        if "dram_read_bytes" in k and "dram_write_bytes" in k:
            k["dram_bytes_per_launch"] = k["dram_read_bytes"] + k["dram_write_bytes"]
        kernels.append(k)
    doc = {"source": rep, "note": note, "kernels": kernels}
    if len(kernels) >= 1 and "dram_bytes_per_launch" in kernels[0]:
        doc["dram_bytes_per_launch"] = max(k.get("dram_bytes_per_launch", 0) for k in kernels)
    with open(out, "w") as f:
        json.dump(doc, f, indent=1)
    for k in kernels:
        print(k["name"][:60], k.get("duration_us"), k.get("tensor_pipe_active_pct"), k.get("alu_pipe_pct"))


if __name__ == "__main__":
    main()
