"""ncu --page raw --csv  ->  a small JSON summary per kernel launch (duration, DRAM bytes, LSU / L2 / issue utilisation,
registers) + the per-GraphSum DRAM traffic bench.py quotes as `roofline.traffic`.

  ncu -i gpurun_out/x.ncu-rep --page raw --csv > profiles/x_ncu_full_raw.csv
  python scripts/summarize_ncu.py profiles/x_ncu_full_raw.csv profiles/graphsum_d16_bittile_summary.json "what was captured"
"""
import csv
import json
import sys

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}
KEYS = {"gpu__time_duration.sum": "gpu_time_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct_of_peak_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_throughput_pct_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_slots_pct_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
        "launch__registers_per_thread": "registers", "sm__inst_executed.sum": "warp_instructions"}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = {"what": sys.argv[3] if len(sys.argv) > 3 else "", "source": sys.argv[1], "kernels": []}
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name", "").split("(")[0].replace("void ", "").strip()}
        for src, dst in KEYS.items():
            if src in d and d[src] not in ("", "n/a"):
                v = float(d[src].replace(",", ""))
                k[dst] = v * UNIT.get(u.get(src, ""), 1.0)
        out["kernels"].append(k)
    out["dram_bytes_per_launch"] = sum(k.get("dram_read_bytes", 0) + k.get("dram_write_bytes", 0) for k in out["kernels"])
    out["gpu_time_us_serialised"] = sum(k.get("gpu_time_us", 0) for k in out["kernels"])
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
