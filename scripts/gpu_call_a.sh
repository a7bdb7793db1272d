#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r02_c4
timeout 300 python scripts/attn_bench.py > ${P}_attn_bench.txt 2>&1; echo "exit $?" >> ${P}_attn_bench.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "fused_forward or training_path or row_sums or legacy_fp32 or fused_tail or statistics_free or channels_last_half" > ${P}_new_tests.log 2>&1; echo "exit $?" >> ${P}_new_tests.log
timeout 1200 python -m pytest tests/test_gpu_reference.py tests/test_gpu_fullsize.py -m gpu -q -rA -s > ${P}_ref_full_tests.log 2>&1; echo "exit $?" >> ${P}_ref_full_tests.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "not (fused_forward or training_path or row_sums or legacy_fp32 or fused_tail or statistics_free or channels_last_half)" > ${P}_rest_tests.log 2>&1; echo "exit $?" >> ${P}_rest_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "exit $?" >> ${P}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "exit $?" >> ${P}_bench.err
for f in ${P}_*tests.log ${P}_smoke.log; do echo "== $f"; tail -n 3 $f; done
cat ${P}_attn_bench.txt
