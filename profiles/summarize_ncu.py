"""Condense `ncu --page raw --csv` output (one row per profiled launch) into
profiles/ncu_r01_summary.json and the per-pass DRAM traffic table
profiles/ncu_traffic.json that bench.py reports as roofline.traffic.

  python profiles/summarize_ncu.py gpurun_out/ncu_r01_raw.csv
"""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_allocated", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "smsp__inst_executed.sum",
]


def scale(value, unit):
    """ncu prints byte counts in scaled units; return bytes."""
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value) * mult.get(unit, 1)


def main(path):
    rows = list(csv.reader(open(path, newline="")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[start], rows[start + 1]
    unit = dict(zip(hdr, units))
    out, traffic = [], {}
    for vals in rows[start + 2:]:
        if len(vals) != len(hdr):
            continue
        d = dict(zip(hdr, vals))
        name = d.get("Kernel Name", "")
        rec = {"kernel": name[:60]}
        for k in KEEP:
            if k in d:
                rec[k] = d[k]
                if unit.get(k):
                    rec[k + " [unit]"] = unit[k]
        stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v)
                  for k, v in d.items()
                  if k.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in k
                  and v not in ("", "n/a")}
        tot = sum(stalls.values()) or 1.0
        rec["stall_pct"] = {k: round(100 * v / tot, 1)
                            for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]}
        out.append(rec)
        if "fast_conv_rows" in name:
            key = "z convolve"
        elif "forward_many<3" in name or "tma_forward_real" in name:
            key = "x forward"
        elif "backward_many<3" in name or "tma_backward_real" in name:
            key = "x backward"
        elif "forward_many<0" in name or "tma_forward_direct" in name:
            key = "y forward"
        elif "backward_many<0" in name or "tma_backward_direct" in name:
            key = "y backward"
        else:
            continue
        b = (scale(d["dram__bytes_read.sum"], unit.get("dram__bytes_read.sum", "byte"))
             + scale(d["dram__bytes_write.sum"], unit.get("dram__bytes_write.sum", "byte")))
        traffic.setdefault(key, []).append(b)
    tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
    with open(os.path.join(HERE, "ncu_%s_summary.json" % tag), "w") as fh:
        json.dump(out, fh, indent=1)
    with open(os.path.join(HERE, "ncu_traffic.json"), "w") as fh:
        json.dump({"source": "profiles/ncu_" + tag + "_summary.json (ncu --set full, one convolution "
                             "of bench.py, B200)",
                   "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}},
                  fh, indent=1)
    for r in out:
        print(r["kernel"][:48], r.get("gpu__time_duration.sum"), r["stall_pct"])


if __name__ == "__main__":
    main(sys.argv[1])
