set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r01b_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python scripts/attn_cl_shape.py 16 32 44 > gpurun_out/plain_attn.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"gemm_kernel|diag_kernel" -s 8 -c 4 -o gpurun_out/r01b_attn_cl python scripts/attn_cl_shape.py 16 32 44 > gpurun_out/ncu_attn.log 2>&1
python scripts/train_kernels_shape.py > gpurun_out/plain_trk.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"bn_|tail_loss" -s 8 -c 4 -o gpurun_out/r01b_train_kernels python scripts/train_kernels_shape.py > gpurun_out/ncu_trk.log 2>&1
tail -2 gpurun_out/ncu_bench.log gpurun_out/ncu_attn.log gpurun_out/ncu_trk.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r01b_launches_bench.csv
