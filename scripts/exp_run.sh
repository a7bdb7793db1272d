set -x
python scripts/bn_time.py
python -m pytest tests -m gpu -x -q -k batchnorm 2>&1 | tail -3
python bench.py --workload train --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_train_graph.json
head -c 400 gpurun_out/bench_train_graph.json; echo
