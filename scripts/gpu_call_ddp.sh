#!/bin/bash
# 2-GPU training-step variants: split / overlapped all-reduce vs single all-reduce; NCCL topology line
mkdir -p gpurun_out
P=gpurun_out/r02_ddp
run() { # name, extra env, extra args
  name=$1; shift
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 20 --warmup 5 ${EXTRA} > ${P}_${name}.json 2> ${P}_${name}.err
  echo "$name exit $?"; tail -c 600 ${P}_${name}.json | head -c 300; echo
}
EXTRA="" run pipelined NCCL_DEBUG=WARN
EXTRA="--pieces 1" run single NCCL_DEBUG=WARN
EXTRA="--split-backward" run split NCCL_DEBUG=WARN
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_ddp_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'img/s %.1f ms %.3f 1s-ms %.3f ar_ms %s early %s sync %s' % (d['value'], d['ms_per_step'], d['timed_1s']['ms_per_step'], d['allreduce_ms'], d['allreduce_early_segment_ms'], d['replicas_in_sync']))
    except Exception as e:
        print(f, 'ERR', e)
PY
