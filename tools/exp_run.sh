#!/bin/bash
# dev tool (GPU box): time env.step for every exp/lib_*.so (driver-style 20-step window and a 160-step window)
for l in exp/lib_*.so; do
  tag=$(basename $l .so)
  for cfg in "20 5" "160 8"; do
    set -- $cfg
    MHPPO_LIB=$PWD/$l python bench.py --no-cpu-baseline --ppo-iters 0 --steps $1 --warmup $2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$tag steps=$1', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
  done
done
