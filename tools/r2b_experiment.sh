#!/bin/bash
# one GPU call: band tests, then the band path on one GPU (self-neighbour) with the early unpack on and off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_svgf.py -x -q -k "band" > gpurun_out/exp4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp4_pytest.log
tail -3 gpurun_out/exp4_pytest.log
: > gpurun_out/exp4_band.txt
for eu in 0 1 0 1; do
  RMD_BAND_EARLY_UNPACK=$eu timeout 200 python tools/band_probe.py --frames 3 --steps 40 --no-plain >> gpurun_out/exp4_band.txt 2>> gpurun_out/exp4_band.err
done
cat gpurun_out/exp4_band.txt
tail -3 gpurun_out/exp4_band.err
