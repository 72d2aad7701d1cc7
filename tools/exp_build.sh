#!/bin/bash
# dev tool: build an experimental variant of the library: tools/exp_build.sh <tag> [-DMACRO=..]...
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p exp/obj_$tag
F="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -fmad=false"
for f in env_inst_scalable mhppo_api; do
  nvcc $F "$@" -Xptxas -v -c mh-ppo_b200/csrc/$f.cu -o exp/obj_$tag/$f.o 2> exp/obj_$tag/$f.log &
done
wait
objs=$(ls mh-ppo_b200/build/*.o | grep -v -e env_inst_scalable.o -e mhppo_api.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o exp/lib_$tag.so $objs exp/obj_$tag/env_inst_scalable.o exp/obj_$tag/mhppo_api.o
grep -A2 "k_env_stepILi5ELi4ELi3" exp/obj_$tag/env_inst_scalable.log | grep -E "Used|spill" | tr '\n' ' '; echo " [$tag]"
