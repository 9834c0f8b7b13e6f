#!/bin/bash
# one GPU call: band tests, then the band path on one GPU (self-neighbour) with the edge launch on the push stream /
# serpentine on and off, then the 128-register ticket (variant 18) against 16
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_svgf.py -x -q -k "band or variants or bit_identical" > gpurun_out/exp3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp3_pytest.log
tail -3 gpurun_out/exp3_pytest.log
: > gpurun_out/exp3_band.txt
for cfg in "0 0" "1 0" "0 1" "1 1" "0 0" "1 1"; do
  set -- $cfg
  RMD_BAND_EDGE_STREAM=$1 RMD_BAND_SERPENTINE=$2 timeout 200 python tools/band_probe.py --frames 3 --steps 40 --no-plain >> gpurun_out/exp3_band.txt 2>> gpurun_out/exp3_band.err
done
RMD_BAND_EDGE_STREAM=1 RMD_BAND_SERPENTINE=1 timeout 200 python tools/band_probe.py --ranks 2 --frames 3 --steps 20 --no-plain >> gpurun_out/exp3_band.txt 2>> gpurun_out/exp3_band.err
RMD_BAND_EDGE_STREAM=0 RMD_BAND_SERPENTINE=0 timeout 200 python tools/band_probe.py --ranks 2 --frames 3 --steps 20 --no-plain >> gpurun_out/exp3_band.txt 2>> gpurun_out/exp3_band.err
cat gpurun_out/exp3_band.txt
timeout 300 python tools/variant_bench.py --workload 4k --frames 12 --steps 36 --pdl 1 --variants 16,18,16,18 > gpurun_out/exp3_variants_4k.jsonl 2> gpurun_out/exp3_variants_4k.err
cut -c1-230 gpurun_out/exp3_variants_4k.jsonl
