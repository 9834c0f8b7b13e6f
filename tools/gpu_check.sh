#!/bin/bash
# final check of the tree: full GPU suite + smoke
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/final2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final2_pytest.log
tail -3 gpurun_out/final2_pytest.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2_smoke.log 2>&1; tail -1 gpurun_out/final2_smoke.log
