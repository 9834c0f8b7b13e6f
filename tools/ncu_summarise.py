"""Summarise an ncu report (`ncu --set full ... -o X`) into the two files the bench and the docs read:
    python tools/ncu_summarise.py gpurun_out/X.ncu-rep profiles/rN_final_metrics.txt [profiles/rN_traffic.json [workload]]
Per kernel name: the LAST captured launch's headline metrics; the traffic file carries DRAM bytes and warp
instructions per launch (a-trous levels averaged) for bench.py's `roofline.traffic` / `second_ceiling`."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rep, out_txt = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    workload = sys.argv[4] if len(sys.argv) > 4 else "1080p"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(head)}
    last = {}
    for r in body:
        last[r[col["Kernel Name"]]] = r
    lines = [f"# {rep}: ncu --set full --clock-control none, last captured launch per kernel"]
    per_kernel, inst = {}, {}
    for name, r in last.items():
        if not any(s in name for s in ("temporal", "variance", "atrous", "box_", "remodulate", "gbuffer_convert")):
            continue
        lines.append(f"kernel: {name}")
        for m in METRICS:
            if m in col:
                lines.append(f"  {m:70s} {r[col[m]]} {units[col[m]]}")

        def nbytes(m):
            return float(r[col[m]]) * UNIT_SCALE.get(units[col[m]], 1.0)
        per_kernel[name] = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
        inst[name] = float(r[col["smsp__inst_executed.sum"]])
    open(out_txt, "w").write("\n".join(lines) + "\n")
    if out_json:
        lv = [k for k in per_kernel if "atrous_kernel" in k]
        json.dump({"workload": workload, "source": out_txt + " (ncu --set full, cold cache, per launch)",
                   "atrous_level_dram_bytes_per_launch": sum(per_kernel[k] for k in lv) / max(len(lv), 1),
                   "atrous_level_warp_inst_per_launch": sum(inst[k] for k in lv) / max(len(lv), 1),
                   "per_kernel": per_kernel, "warp_inst": inst}, open(out_json, "w"), indent=1)
    print("\n".join(lines[:8]))


if __name__ == "__main__":
    main()
