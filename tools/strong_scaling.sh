#!/bin/bash
# Strong scaling of BASELINE configs 4 and 5 on 1 and 2 GPUs of one box (fixed global batch), one RESULT line per run:
#   gpurun --gpus 2 -- bash tools/strong_scaling.sh
P='import json,sys
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print("RESULT", d["n_gpus"], d["config"]["workload"][:50], d["config"]["graphs_per_gpu"], d["value"], d["ms_per_step"], d["median_ms_per_step"], d["gpu_launches"])'
timeout 300 python bench.py --config affinity_ecfp --no-cpu-baseline --no-roofline --steps 40 2>&1 | python -c "$P"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --config affinity_ecfp --batch 1024 --no-cpu-baseline --no-roofline --steps 40 2>&1 | python -c "$P"
timeout 300 python bench.py --config autoenc --batch 8192 --hidden 64 --no-cpu-baseline --no-roofline --steps 30 2>&1 | python -c "$P"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --config autoenc --batch 4096 --hidden 64 --no-cpu-baseline --no-roofline --steps 30 2>&1 | python -c "$P"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 2 --no-cpu-baseline --no-roofline --steps 60 2>&1 | python -c "$P"
