#!/bin/bash
# Run on the GPU box, one ncu pass per call:
#   tools/gpu_profile.sh <tag> list    plain bench, then the ncu launch list of the same command
#   tools/gpu_profile.sh <tag> full    one full capture of the two top kernels of the frame
#   tools/gpu_profile.sh <tag> train   one full capture of the training-step kernels
set -u
TAG=${1:-r1}
MODE=${2:-list}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --views 8 --no-cpu-baseline --no-train"
case $MODE in
list)
  python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
  tail -1 gpurun_out/bench_$TAG.json
  $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_list_$TAG.log ;;
full)
  $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'trace_compact_kernel|ngp_forward' -s 8 -c 4 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_full_$TAG.log ;;
extra)
  # the other kernels of the north star: baked shading + K=32 traversal on configs[4] (4K, 1.15 M triangles)
  XCMD="python tools/run_leg.py c5 3"
  $XCMD > gpurun_out/plain_extra_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'baked_shade_kernel|trace_compact_kernel|composite_rays_kernel' -s 12 -c 3 -o gpurun_out/prof_extra_$TAG $XCMD > gpurun_out/ncu_extra_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_extra_$TAG.log ;;
misc)
  # the remaining kernels of the path: segmented-scan compositing, the quadrature Field net, the occupancy-grid marcher
  MCMD="python tools/run_field_legs.py 5"
  $MCMD > gpurun_out/plain_misc_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'render_weights|field_net_forward|field_net_backward|occgrid_march|accumulate_|derive_properties' -s 40 -c 14 -o gpurun_out/prof_misc_$TAG $MCMD > gpurun_out/ncu_misc_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_misc_$TAG.log
  ncu --set full --clock-control none --import-source on -k regex:'occgrid_march' -s 6 -c 1 -o gpurun_out/prof_misc2_$TAG $MCMD > gpurun_out/ncu_misc2_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_misc2_$TAG.log ;;
train)
  TCMD="python tools/diag_train.py 3"
  $TCMD > gpurun_out/plain_train_$TAG.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'ngp_backward_kernel|weight_grad_kernel|trace_refill_kernel' -s 9 -c 3 -o gpurun_out/prof_train_$TAG $TCMD > gpurun_out/ncu_train_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_train_$TAG.log ;;
esac
ls -la gpurun_out/ | tail -12
