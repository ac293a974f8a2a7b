"""Summarise ncu output for profiles/:
    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_summary.md
    python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_k2_full.json [chain_steps]
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        val = float(r[-1].replace(",", ""))
        unit = r[-2]
        ms = val * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(name, [0, 0.0, r[7], r[8]])
        a[0] += 1; a[1] += ms
        total += ms
    with open(out, "w") as f:
        f.write("# ncu launch list summary (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write(f"source: `{path}`; {len(rows)} launches, {total:.3f} ms total\n\n| kernel | launches | total ms | share | block | grid |\n|---|---|---|---|---|---|\n")
        for name, (n, ms, blk, grd) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% | {blk} | {grd} |\n")
    print(open(out).read())


WANT = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_blocks", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def full(path, out, chain_steps=None):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, unit = rows[0], rows[1]
    res = []
    for val in rows[2:]:
        d = OrderedDict()
        d["kernel"] = val[hdr.index("Kernel Name")]
        for h, u, v in zip(hdr, unit, val):
            if h in WANT or h.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in h:
                try:
                    d[h] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    pass
        if chain_steps and "smsp__inst_executed.sum" in d:
            d["warp_instructions_per_chain_step"] = d["smsp__inst_executed.sum"]["value"] / chain_steps
        rd = d.get("dram__bytes_read.sum"); wr = d.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            d["dram_bytes_per_launch"] = rd["value"] * scale.get(rd["unit"], 1) + wr["value"] * scale.get(wr["unit"], 1)
        res.append(d)
    json.dump(res if len(res) > 1 else res[0], open(out, "w"), indent=1)
    print(json.dumps(res[0], indent=1)[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], float(sys.argv[4]) if len(sys.argv) > 4 else None)
