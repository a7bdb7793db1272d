#!/bin/bash
# 2 GPUs: BN tests + micro-benchmark, then the training step pipelined (default) vs split backward with BN grids that leave NCCL's SMs free
mkdir -p gpurun_out
P=gpurun_out/r02_f
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "batchnorm or bn" > ${P}_bn_tests.log 2>&1; echo "exit $?" >> ${P}_bn_tests.log
timeout 300 python scripts/bn_time.py > ${P}_bn_time.log 2>&1; echo "exit $?" >> ${P}_bn_time.log
run() {  # name port extra-args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --workload train --steps 20 --warmup 5 ${@:3} 2> ${P}_$1.err | tail -1 > ${P}_$1.json
  echo "$1 exit $?"
}
timeout 240 python bench.py --workload train --steps 20 --warmup 5 2> ${P}_n1.err | tail -1 > ${P}_n1.json
run pipelined 29611
run split16 29612 --split-backward --nccl-sms 16
run split8 29613 --split-backward --nccl-sms 8
run split32 29614 --split-backward --nccl-sms 32
tail -n 3 ${P}_bn_tests.log; head -n 9 ${P}_bn_time.log
python - <<PY
import json
for f in ("n1","pipelined","split16","split8","split32"):
    try:
        d=json.load(open(f"gpurun_out/r02_f_{f}.json"))
        print(f, round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms", d.get("replicas_in_sync"), d.get("clocks",{}).get("sm_mhz"))
    except Exception as e:
        print(f, "missing/failed", e)
PY
