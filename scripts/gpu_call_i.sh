#!/bin/bash
# 1 GPU: split-schedule test; 2 GPUs: three-graph overlapped step, timeline + bench A/B
mkdir -p gpurun_out
P=gpurun_out/r02_i
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "split_backward or graphed_train or sm_limit" > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
timeout 240 python bench.py --workload train --steps 20 --warmup 5 2> ${P}_n1.err | tail -1 > ${P}_n1.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 scripts/ddp_overlap_probe.py > ${P}_probe16.log 2>&1; echo "exit $?" >> ${P}_probe16.log
run() {  # name port extra-args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --workload train --steps 20 --warmup 5 ${@:3} 2> ${P}_$1.err | tail -1 > ${P}_$1.json
  echo "$1 exit $?"
}
run pipelined 29642
run split16 29643 --split-backward --nccl-sms 16
run split12 29644 --split-backward --nccl-sms 12
tail -n 3 ${P}_tests.log; grep -v Warning ${P}_probe16.log | grep "ms\|exit" | tail -n 10
python - <<PY
import json
for f in ("n1","pipelined","split16","split12"):
    try:
        d=json.load(open(f"gpurun_out/r02_i_{f}.json"))
        print(f, round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms", d.get("replicas_in_sync"), d.get("clocks",{}).get("sm_mhz"))
    except Exception as e:
        print(f, "missing/failed", e)
PY
