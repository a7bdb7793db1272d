#!/bin/bash
# 2 GPUs: whole GPU suite + smoke on one of them, then the DEFAULT bench at N = 1 and N = 2 exactly as the driver launches it
mkdir -p gpurun_out
P=gpurun_out/r02_j
timeout 2700 python -m pytest tests -m gpu -q -rA -s > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "exit $?" >> ${P}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench_n1.json 2> ${P}_bench_n1.err; echo "exit $?" >> ${P}_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 20 --warmup 5 > ${P}_bench_n2.json 2> ${P}_bench_n2.err; echo "exit $?" >> ${P}_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > ${P}_bench_ref_n2.json 2> ${P}_bench_ref_n2.err; echo "exit $?" >> ${P}_bench_ref_n2.err
grep -E "passed|failed|error" ${P}_tests.log | tail -n 5
tail -n 3 ${P}_smoke.log
python - <<PY
import json
for f in ("bench_n1","bench_n2","bench_ref_n2"):
    try:
        d=json.loads(open(f"gpurun_out/r02_j_{f}.json").read().strip().splitlines()[-1])
        tr=d.get("train") or {}
        print(f, round(d["value"],1), d["unit"], "e2e", (d.get("e2e") or {}).get("value"), "| train", tr.get("value"), tr.get("ms_per_step"), tr.get("replicas_in_sync"), tr.get("allreduce_ms"), tr.get("allreduce_exposed_segment_ms"))
    except Exception as e:
        print(f, "missing/failed", e)
PY
