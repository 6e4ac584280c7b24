#!/bin/bash
# Run on the GPU box: plain bench, then ncu launch list, then one full capture of the two top kernels.
# usage: tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
tail -1 gpurun_out/bench_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --views 8 --no-cpu-baseline --no-train"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'trace_compact_kernel|ngp_forward_kernel' -s 8 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
# training step kernels (backward, weight-gradient GEMM, incoherent-ray trace)
TCMD="python tools/diag_train.py 3"
$TCMD > gpurun_out/plain_train_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'ngp_backward_kernel|weight_grad_kernel|trace_kernel' -s 9 -c 3 -o gpurun_out/prof_train_$TAG $TCMD > gpurun_out/ncu_train_$TAG.log 2>&1
tail -2 gpurun_out/ncu_train_$TAG.log
ls -la gpurun_out/
