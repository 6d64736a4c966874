#!/bin/bash
# Round-2 evidence run on ONE B200 (under gpurun): GPU tests, the default bench line, then ncu launch lists and
# --set full captures of the dominant kernels -- each ncu command only after the same command exited 0 without it.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_gputests.log 2>&1; tail -3 gpurun_out/r2_final_gputests.log
python bench.py > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err || { tail -5 gpurun_out/r2_final_bench_1gpu.err; }
B="python bench.py --nmol 20000 --steps 2 --warmup 3 --no-extra --cpu-sample 500"
$B > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_bench_nmol20000.csv $B > gpurun_out/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eeq_factor_ll -s 5 -c 1 -o gpurun_out/r2_factor $B > gpurun_out/r2_ncu_factor.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eeq_batch_kernel -s 10 -c 2 -o gpurun_out/r2_pair $B > gpurun_out/r2_ncu_pair.log 2>&1
NO_DQ=1 python tools/time_cluster.py 16384 > gpurun_out/r2_cluster_plain.log 2>&1 || exit 1
NO_DQ=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_cluster16k_grad.csv python tools/time_cluster.py 16384 > gpurun_out/r2_ncu_cluster.log 2>&1
python tools/prof_gemm.py 15360 512 > gpurun_out/r2_gemm_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 1 -c 1 -o gpurun_out/r2_gemm python tools/prof_gemm.py 15360 512 > gpurun_out/r2_ncu_gemm.log 2>&1
ls -la gpurun_out/*.ncu-rep
