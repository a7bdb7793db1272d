# usage: bash scripts/sweep.sh N   -> gpurun_out/sweep_{infer,train}_nN.json
N=$1
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --workload train --steps 20 --warmup 5 2>gpurun_out/sweep_train_n$N.err | tail -1 > gpurun_out/sweep_train_n$N.json
if [ "$2" = "both" ]; then
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/sweep_infer_n$N.err | tail -1 > gpurun_out/sweep_infer_n$N.json
fi
python - <<PY
import json
for f in ("sweep_train_n$N","sweep_infer_n$N"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"],1), d.get("n_gpus"), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d.get("replicas_in_sync"), d.get("clocks"))
    except Exception as e:
        print(f, "missing/failed", e)
PY
