#!/bin/bash
# 8 GPUs: the DEFAULT bench exactly as the driver launches it, then the training step with the pipelined schedule for the A/B
mkdir -p gpurun_out
P=gpurun_out/r02_k
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_tensor or graphed_train" > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 8 --steps 20 --warmup 5 > ${P}_bench_n8.json 2> ${P}_bench_n8.err; echo "exit $?" >> ${P}_bench_n8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29662 bench.py --gpus 8 --workload train --no-split-backward --steps 20 --warmup 5 > ${P}_train_pipelined_n8.json 2> ${P}_train_pipelined_n8.err; echo "exit $?" >> ${P}_train_pipelined_n8.err
timeout 300 python bench.py --workload train --steps 20 --warmup 5 > ${P}_train_n1.json 2> ${P}_train_n1.err
tail -n 2 ${P}_tests.log
python - <<PY
import json
for f in ("bench_n8","train_pipelined_n8","train_n1"):
    try:
        d=json.loads(open(f"gpurun_out/r02_k_{f}.json").read().strip().splitlines()[-1])
        tr=d.get("train") or (d if "train" in f else {})
        print(f, round(d["value"],1), d["unit"], "e2e", (d.get("e2e") or {}).get("value"), "| train", tr.get("value"), tr.get("ms_per_step"), tr.get("replicas_in_sync"), tr.get("allreduce_ms"), tr.get("allreduce_exposed_segment_ms"))
    except Exception as e:
        print(f, "missing/failed", e)
PY
