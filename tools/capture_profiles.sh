#!/bin/bash
# ncu evidence for one build (run on a GPU box: bash tools/capture_profiles.sh <tag>): the launch list of the bench command
# and --set full captures of the convolution family and of the HBM-bound kernels. Summaries: tools/summarize_profiles.py.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launches.log 2>&1
FULL="timeout 900 ncu --set full --clock-control none --import-source on -f"
$FULL -k regex:conv3x3_tc_kernel -s 0 -c 4 -o gpurun_out/${TAG}_conv_fwd16 $B > gpurun_out/${TAG}_ncu_a.log 2>&1
$FULL -k regex:conv3x3_tc_kernel -s 8 -c 6 -o gpurun_out/${TAG}_conv_f16 $B > gpurun_out/${TAG}_ncu_b.log 2>&1
$FULL -k regex:conv3x3_wgrad_kernel -s 2 -c 2 -o gpurun_out/${TAG}_wgrad_f16 $B > gpurun_out/${TAG}_ncu_c.log 2>&1
$FULL -k "regex:gn_bwd_apply|gn_apply_kernel|in_mse|moments|boxsum_kernel|paint|nchw_to_nhwc|nhwc_to_nchw" -c 30 -o gpurun_out/${TAG}_hbm $B > gpurun_out/${TAG}_ncu_d.log 2>&1
timeout 600 $FULL -k regex:token_program -c 2 -o gpurun_out/${TAG}_tokenprog $B > gpurun_out/${TAG}_ncu_e.log 2>&1
# gpurun copies back at most 64 MiB: keep the raw metric tables of every capture, and the reports (with source) only of the
# convolution kernels
for r in conv_fwd16 conv_f16 wgrad_f16 hbm tokenprog; do
  if [ -f gpurun_out/${TAG}_$r.ncu-rep ]; then
    ncu -i gpurun_out/${TAG}_$r.ncu-rep --page raw --csv > gpurun_out/${TAG}_${r}_raw.csv 2>/dev/null
  fi
done
rm -f gpurun_out/${TAG}_hbm.ncu-rep gpurun_out/${TAG}_tokenprog.ncu-rep gpurun_out/${TAG}_conv_fwd16.ncu-rep
ls -la gpurun_out/${TAG}_*
