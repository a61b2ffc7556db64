#!/bin/bash
# Launch-time experiments on the tensor-core context kernel (BASIC_TC_DEBUG bits: 1 = A producers skip loads,
# 2 = no weight copies, 4 = no MMAs, 8 = drain skips tcgen05.ld, 16 = producers skip split + store).
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_layer_tc -c 16 --csv \
    --log-file gpurun_out/exp_$name.csv python tools/profile_step.py cfg2 1 > /dev/null 2>&1
  python - gpurun_out/exp_$name.csv "$name" <<'P'
import csv,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
h=rows[0]; vi=h.index("Metric Value")
print(sys.argv[2],[round(float(r[vi])/1000,1) for r in rows[1:]][8:16])
P
}
for spec in "$@"; do
  run "$(echo $spec | tr ' =' '__')" $spec
done
