#!/bin/bash
# ncu evidence for profiles/: (1) launch list of the bench's eager steps, (2) full captures of the hand-written kernels
mkdir -p gpurun_out
P=gpurun_out/r02_ncu
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-train --no-graph --no-long"
$CMD > ${P}_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "acan_step/" -c 600 --csv --log-file ${P}_launches_bench.csv $CMD > ${P}_ncu_bench.log 2>&1
echo "launch list exit $?"
python scripts/ncu_targets.py > ${P}_plain_targets.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "ncu_capture/" -o ${P}_kernels python scripts/ncu_targets.py > ${P}_ncu_targets.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/r02_ncu*
tail -n 3 ${P}_ncu_bench.log ${P}_ncu_targets.log
