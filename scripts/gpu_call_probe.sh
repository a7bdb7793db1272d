#!/bin/bash
mkdir -p gpurun_out
P=gpurun_out/r02_probe
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 scripts/ddp_overlap_probe.py > ${P}_default.txt 2>&1; echo "exit $?" >> ${P}_default.txt
NCCL_CGA_CLUSTER_SIZE=0 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 scripts/ddp_overlap_probe.py > ${P}_cga0.txt 2>&1; echo "exit $?" >> ${P}_cga0.txt
NCCL_MAX_CTAS=8 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 scripts/ddp_overlap_probe.py > ${P}_ctas8.txt 2>&1; echo "exit $?" >> ${P}_ctas8.txt
grep -h "ms from\|late all\|exit" ${P}_default.txt ${P}_cga0.txt ${P}_ctas8.txt
