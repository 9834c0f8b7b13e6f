#!/bin/bash
# one GPU call: the variant / variance-path bit-identity tests, then A/B of variant 17 and the reversed variance walk
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_svgf.py -x -q -k "bit_identical or variants or band" > gpurun_out/exp2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/exp2_pytest.log
tail -3 gpurun_out/exp2_pytest.log
timeout 300 python tools/variant_bench.py --workload 4k --frames 12 --steps 36 --pdl 1 \
  --variants 6,17,17+RMD_VAR_REVERSE=0,16@0,17,6 > gpurun_out/exp2_variants_4k.jsonl 2> gpurun_out/exp2_variants_4k.err
timeout 300 python tools/variant_bench.py --workload 8k --frames 3 --steps 10 --warmup 4 --pdl 1 \
  --variants 6,17,6,17 > gpurun_out/exp2_variants_8k.jsonl 2> gpurun_out/exp2_variants_8k.err
timeout 200 python tools/variant_bench.py --workload 1080p --frames 12 --steps 48 --pdl 1 \
  --variants 6,17,17+RMD_VAR_REVERSE=0,6 > gpurun_out/exp2_variants_1080p.jsonl 2> gpurun_out/exp2_variants_1080p.err
cut -c1-230 gpurun_out/exp2_variants_*.jsonl
