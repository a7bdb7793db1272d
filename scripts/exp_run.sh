set -x
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload train --steps 10 --warmup 3 2>&1 | tail -4 > gpurun_out/bench_train_n2.log
tail -c 1800 gpurun_out/bench_train_n2.log
