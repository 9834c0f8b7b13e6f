#!/bin/bash
# ncu evidence for the final round-2 kernels: plain run first, then the launch list and one --set full capture (4K frame 4),
# then the five level launches of an 8K frame
mkdir -p gpurun_out
timeout 200 python tools/profile_frame.py --workload 4k --frames 5 > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2b_launches_4k.csv \
  python tools/profile_frame.py --workload 4k --frames 6 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:"atrous|temporal|variance" -s 28 -c 7 -f -o gpurun_out/r2b_4k \
  python tools/profile_frame.py --workload 4k --frames 5 > gpurun_out/ncu_full_4k.log 2>&1; echo "full 4k rc=$?"
timeout 400 ncu --set full --clock-control none -k regex:"atrous" -s 15 -c 5 -f -o gpurun_out/r2b_8k \
  python tools/profile_frame.py --workload 8k --frames 4 > gpurun_out/ncu_full_8k.log 2>&1; echo "full 8k rc=$?"
ls -la gpurun_out/*.ncu-rep
