#!/bin/bash
# 2 GPUs: split-step timeline (with / without the collective), NCCL channel count of the exposed all-reduce; 1 GPU: BN + tail-loss tests, train bench
mkdir -p gpurun_out
P=gpurun_out/r02_h
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "batchnorm or bn or tail or ordinal" > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
timeout 240 python bench.py --workload train --steps 20 --warmup 5 2> ${P}_n1.err | tail -1 > ${P}_n1.json
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 scripts/ddp_overlap_probe.py > ${P}_probe16.log 2>&1; echo "exit $?" >> ${P}_probe16.log
run() {  # name port extra-args
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --workload train --steps 20 --warmup 5 ${@:3} 2> ${P}_$1.err | tail -1 > ${P}_$1.json
  echo "$1 exit $?"
}
run pipelined 29632
run split16 29633 --split-backward --nccl-sms 16
run split8 29634 --split-backward --nccl-sms 8
run minctas32 29635 --nccl-min-ctas 32
tail -n 3 ${P}_tests.log; grep -v Warning ${P}_probe16.log | grep "ms\|exit" | tail -n 12
python - <<PY
import json
for f in ("n1","pipelined","split16","split8","minctas32"):
    try:
        d=json.load(open(f"gpurun_out/r02_h_{f}.json"))
        print(f, round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms", d.get("replicas_in_sync"), d.get("clocks",{}).get("sm_mhz"))
    except Exception as e:
        print(f, "missing/failed", e)
PY
