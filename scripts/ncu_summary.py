"""Summarise an .ncu-rep (ncu --set full) into a small JSON: the numbers DESIGN.md and
bench.py's roofline.traffic cite.  Usage: python scripts/ncu_summary.py REP OUT.json [--kernel regex]"""
import csv
import json
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
    "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_src_fp64.sum", "sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.avg.per_cycle_elapsed",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[3] == "--kernel" else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if pat and not pat.search(name):
            continue
        d = {"kernel": name}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if u in SCALE and SCALE[u] != 1.0:
                    v, u = v * SCALE[u], ("byte" if "byte" in u else "s")
                d[k] = {"value": v, "unit": u}
        stalls = {}
        for i, h in enumerate(hdr):
            m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h)
            if m and not h.endswith("_not_issued"):
                try:
                    stalls[m.group(1)] = float(r[i].replace(",", ""))
                except ValueError:
                    pass
        tot = sum(stalls.values())
        if tot > 0:
            d["warp_stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]}
        if "dram__bytes_read.sum" in d and "dram__bytes_write.sum" in d:
            d["dram_bytes_per_launch"] = d["dram__bytes_read.sum"]["value"] + d["dram__bytes_write.sum"]["value"]
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1)[:3000])


if __name__ == "__main__":
    main()
