for d in 0 1 2 4 3 7; do
  BASIC_TC_DEBUG=$d ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_layer_tc" --csv --log-file gpurun_out/exp_$d.csv python tools/profile_step.py cfg2 1 > gpurun_out/ncu.log 2>&1
  echo "debug=$d"; python tools/launch_summary.py gpurun_out/exp_$d.csv --seq | grep k_layer_tc | sed -n 6,9p
done
