timeout 300 python -m pytest tests -m gpu -x -q --tb=short -k "graphed_train" 2>&1 | tail -12
timeout 300 python bench.py --workload train --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_train.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_train.json"))
print("train", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["gpu_launches"], d["clocks"], d["loss"])
PY
