#!/bin/bash
# first GPU call of round 2: new training-path kernels, reference-on-GPU tests, full-size parity, then the rest, then benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02_c1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "training_path or row_sums or legacy_fp32" > gpurun_out/r02_c1_train_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_c1_train_tests.log
timeout 1200 python -m pytest tests/test_gpu_reference.py -m gpu -q -rA -s > gpurun_out/r02_c1_ref_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_c1_ref_tests.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -rA -s > gpurun_out/r02_c1_full_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_c1_full_tests.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "not training_path and not row_sums and not legacy_fp32" > gpurun_out/r02_c1_rest_tests.log 2>&1
echo "exit $?" >> gpurun_out/r02_c1_rest_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c1_smoke.log 2>&1
echo "exit $?" >> gpurun_out/r02_c1_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c1_bench_infer.json 2> gpurun_out/r02_c1_bench_infer.err
timeout 600 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r02_c1_bench_train.json 2> gpurun_out/r02_c1_bench_train.err
tail -5 gpurun_out/r02_c1_*tests.log gpurun_out/r02_c1_smoke.log
