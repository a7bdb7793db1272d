set -x
timeout 400 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -4
timeout 400 python -m pytest tests -m gpu -x -q --tb=short -k "graphed" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_infer.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_infer.json"))
print("infer", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["gpu_launches"], "pv_frac", round(d["roofline"]["frac"],3), "core", round(d["attention_core"]["frac"],3), round(d["attention_core"]["call_us"],1), d["tail"]["launch_us"], d["clocks"])
PY
