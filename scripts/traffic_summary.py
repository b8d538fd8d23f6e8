#!/usr/bin/env python
"""ncu CSV (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over one bench step) -> per-phase
DRAM traffic table (JSON) that bench.py quotes in roofline.traffic.

usage: python scripts/traffic_summary.py gpurun_out/traffic_cfg2.csv profiles/r01k_traffic_cfg2.json
"""
import csv
import json
import sys

PHASE_OF = [("k_drift_key", "drift_key"), ("k_mask_ext", "drift_key"), ("k_drift_count", "drift_count"), ("k_flag_compact", "drift_count"), ("k_vfield_sq", "drift_count"),
            ("k_scan_", "drift_scan"), ("k_drift_place", "drift_place"), ("k_build_ext", "buffer"), ("k_tile_counts", "buffer"),
            ("k_fine_deposit", "fine_deposit"), ("k_fft_x_fwd", "fine_fft_x"), ("k_fft_y<16, 18, 1", "fine_ifft_y"), ("k_fft_y<", "fine_fft_y"),
            ("k_fft_z_green", "fine_fft_z_green"), ("k_fft_x_inv3", "fine_ifft_x"), ("k_fine_kick", "fine_kick"),
            ("k_coarse_deposit", "coarse_deposit"), ("k_coarse_cell_sums", "coarse_deposit"), ("k_coarse_gather27", "coarse_deposit"), ("regular_fft", "coarse_fft_green"), ("vector_fft", "coarse_fft_green"),
            ("k_green", "coarse_fft_green"), ("k_force_c_finish", "coarse_fft_green"), ("k_coarse_kick", "coarse_kick")]


def main(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    col = {n: i for i, n in enumerate(rows[0])}
    per = {}
    for r in rows[1:]:
        per.setdefault(r[col["ID"]], {"kernel": r[col["Kernel Name"]]})[r[col["Metric Name"]]] = float(r[col["Metric Value"]].replace(",", ""))
    phases, kernels = {}, []
    for k in per.values():
        name = k["kernel"].split("(")[0]
        ph = next((p for pat, p in PHASE_OF if pat in k["kernel"]), "other")
        rd, wr = k.get("dram__bytes_read.sum", 0.0), k.get("dram__bytes_write.sum", 0.0)
        kernels.append({"kernel": name, "phase": ph, "dram_read_bytes": rd, "dram_write_bytes": wr, "ns": k.get("gpu__time_duration.sum", 0.0)})
        a = phases.setdefault(ph, {"dram_bytes": 0.0, "launches": 0, "ns": 0.0})
        a["dram_bytes"] += rd + wr; a["launches"] += 1; a["ns"] += k.get("gpu__time_duration.sum", 0.0)
    out = {"source": src, "note": "one PM step under ncu (cold cache, serialised): dram__bytes_read.sum + dram__bytes_write.sum per kernel",
           "total_dram_bytes": sum(p["dram_bytes"] for p in phases.values()), "phases": phases, "kernels": kernels}
    json.dump(out, open(dst, "w"), indent=1)
    for ph, a in phases.items():
        print(f"{ph:20s} {a['dram_bytes'] / 1e9:8.2f} GB  {a['ns'] / 1e6:7.3f} ms  {a['launches']} launches")
    print(f"total {out['total_dram_bytes'] / 1e9:.1f} GB")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
