set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/attn_cl_time.py 16 32 44 20 --kernels
python scripts/attn_cl_time.py 32 20 64 20
python scripts/attn_cl_time.py 16 28 38 20
python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_mid.json
python scripts/train_prof.py amp 45
python scripts/train_prof.py fp32 25
