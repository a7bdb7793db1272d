#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r02_c5
timeout 300 python scripts/attn_bench.py > ${P}_attn_bench.txt 2>&1; echo "exit $?" >> ${P}_attn_bench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "fused_forward or statistics_free or channels_last_half" > ${P}_new_tests.log 2>&1; echo "exit $?" >> ${P}_new_tests.log
tail -n 4 ${P}_new_tests.log
cat ${P}_attn_bench.txt
