#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r02_c7
timeout 300 python scripts/attn_bench.py > ${P}_attn_bench.txt 2>&1; echo "exit $?" >> ${P}_attn_bench.txt
timeout 400 python scripts/train_ops.py > ${P}_train_ops.txt 2>&1; echo "exit $?" >> ${P}_train_ops.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "training_path or optimize_for_inference or graphed_predictor or graphed_train" > ${P}_tests.log 2>&1; echo "exit $?" >> ${P}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras > ${P}_bench.json 2> ${P}_bench.err; echo "exit $?" >> ${P}_bench.err
tail -n 4 ${P}_tests.log
grep -E "train forward|backward|loss C4|per pass:" ${P}_attn_bench.txt
head -75 ${P}_train_ops.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c7_bench.json').read().strip().splitlines()[-1])
print('infer', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], '1s', d['timed_1s']['value'], 'core in-step', d['attention_core']['us_in_step'], d['attention_core']['frac_in_step'], 'pv', d['roofline']['launch_us'], d['roofline']['frac'])
t=d['train']; print('train', t['value'], t['ms_per_step'], 'bwd', t['attention_bwd']['launch_us'], 'fwd', t['attention_fwd_train']['launch_us'])
print('parity', d.get('parity'))
PY
