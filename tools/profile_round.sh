#!/bin/bash
# Profiling pass of a round (B200_PROFILING.md recipe): plain run first, then the ncu launch list of the same command,
# then one `--set full` capture of the dominant kernels.  Outputs under gpurun_out/ (copy summaries to profiles/).
set -x
tag=${1:-r1}
mkdir -p gpurun_out
python tools/profile_step.py cfg2 2 > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py cfg2 2 > gpurun_out/ncu_launches_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.txt
ncu --set full --clock-control none --import-source on -k regex:k_layer_tc -s 8 -c 8 -f -o gpurun_out/prof_${tag}_ctx \
    python tools/profile_step.py cfg2 2 > gpurun_out/ncu_full_ctx_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_bls_|k_pair_|k_quantize|k_dequantize|k_nchw_to_cl' -s 12 -c 14 -f -o gpurun_out/prof_${tag}_coder \
    python tools/profile_step.py cfg2 2 > gpurun_out/ncu_full_coder_$tag.log 2>&1
for f in ctx coder; do
  ncu -i gpurun_out/prof_${tag}_$f.ncu-rep --page raw --csv > gpurun_out/ncu_full_${tag}_$f.csv 2>/dev/null
done
tail -25 gpurun_out/launches_$tag.txt
