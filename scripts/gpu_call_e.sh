#!/bin/bash
# round-2 verification call: whole GPU suite, smoke, BN / attention micro-benchmarks, the default bench line
mkdir -p gpurun_out
P=gpurun_out/r02_e
timeout 2700 python -m pytest tests -m gpu -q -rA -s > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; echo "exit $?" >> ${P}_smoke.log
timeout 300 python scripts/bn_time.py > ${P}_bn_time.log 2>&1; echo "exit $?" >> ${P}_bn_time.log
timeout 300 python scripts/attn_bench.py > ${P}_attn_bench.log 2>&1; echo "exit $?" >> ${P}_attn_bench.log
timeout 900 python bench.py --steps 20 --warmup 5 > ${P}_bench.json 2> ${P}_bench.err; echo "exit $?" >> ${P}_bench.err
grep -E "passed|failed|error" ${P}_tests.log | tail -n 5
tail -n 3 ${P}_smoke.log
head -n 12 ${P}_bn_time.log
grep -E "train|core C2" ${P}_attn_bench.log
head -c 1500 ${P}_bench.json
