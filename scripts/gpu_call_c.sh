#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r02_c6
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "flat_sgd or graphed_train or training_step" > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
tail -n 8 ${P}_tests.log
timeout 600 python bench.py --workload train --steps 20 --warmup 5 > ${P}_train.json 2> ${P}_train.err; echo "exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c6_train.json').read().strip().splitlines()[-1])
print('train img/s %.1f ms %.3f 1s-ms %.3f e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['timed_1s']['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
PY
tail -n 3 ${P}_train.err
