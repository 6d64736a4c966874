#!/bin/bash
# Final evidence run of round 2 on ONE B200 (under gpurun): GPU tests, the default bench line, then the ncu launch list
# and --set full captures of the batch kernels -- each ncu command only after the same command exited 0 without it.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_gputests.log 2>&1; tail -3 gpurun_out/r2f_gputests.log
python bench.py > gpurun_out/r2f_bench_1gpu.json 2> gpurun_out/r2f_bench_1gpu.err || { tail -5 gpurun_out/r2f_bench_1gpu.err; }
B="python bench.py --nmol 20000 --steps 2 --warmup 3 --no-extra --cpu-sample 500"
$B > gpurun_out/r2f_ncu_plain.json 2> gpurun_out/r2f_ncu_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2f_launches_bench_nmol20000.csv $B > gpurun_out/r2f_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eeq_factor_ll -s 3 -c 1 -o gpurun_out/r2f_factor $B > gpurun_out/r2f_ncu_factor.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eeq_batch_kernel -s 6 -c 2 -o gpurun_out/r2f_pair $B > gpurun_out/r2f_ncu_pair.log 2>&1
ls -la gpurun_out/r2f_*.ncu-rep
