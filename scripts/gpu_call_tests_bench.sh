#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r02_c3
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "training_path or row_sums or legacy_fp32" > ${P}_train_tests.log 2>&1; echo "exit $?" >> ${P}_train_tests.log
timeout 1200 python -m pytest tests/test_gpu_reference.py -m gpu -q -rA -s > ${P}_ref_tests.log 2>&1; echo "exit $?" >> ${P}_ref_tests.log
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -rA -s > ${P}_full_tests.log 2>&1; echo "exit $?" >> ${P}_full_tests.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "not training_path and not row_sums and not legacy_fp32" > ${P}_rest_tests.log 2>&1; echo "exit $?" >> ${P}_rest_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "exit $?" >> ${P}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "exit $?" >> ${P}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${P}_bench_ref.json 2> ${P}_bench_ref.err
for f in ${P}_*tests.log ${P}_smoke.log; do echo "== $f"; tail -n 4 $f; done
