#!/bin/bash
# dev tool (GPU box): time env.step for every exp/lib_*.so
for l in exp/lib_*.so; do
  tag=$(basename $l .so)
  MHPPO_LIB=$PWD/$l python bench.py --no-cpu-baseline --ppo-iters 0 --steps 200 --warmup 10 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$tag', round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done
