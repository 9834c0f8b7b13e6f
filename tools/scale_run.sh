#!/bin/bash
# usage: tools/scale_run.sh N  — both sharded modes at N GPUs, compact output
N=$1
P=$((29600 + N))
echo "== $N GPUs: independent 1080p sequences (weak scaling)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 30 --warmup 10 --frames 16 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}), 'e2e', round(d['e2e']['value']))"
echo "== $N GPUs: 8K frame row-banded (strong scaling)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+50)) bench.py --gpus $N --mode banded --workload 8k --steps 20 --warmup 6 --frames 4 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling')}), {k: d['config'].get(k) for k in ('scheme', 'exchange', 'band_rows', 'ext_rows')}, d.get('exchange_bytes_per_frame_per_boundary'))"
