#!/bin/bash
# A/B build of the fused plan kernel: recompiles ONE horizon's translation unit with extra -D flags and links it
# with the already-built objects of the other units into mbpo_b200/ab/libmbpo_<name>.so (git-ignored; travels to the
# GPU box).  Run the bench against it with MBPO_B200_LIB=<that path>.   usage: tools/ab_variant.sh <name> <H> <flags...>
set -e
cd "$(dirname "$0")/.."
name=$1; H=$2; shift 2
pkg=model-based-policy-optimizers_b200
mkdir -p $pkg/build/ab $pkg/mbpo_b200/ab
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I include \
     -DMBPO_INST_H=$H "$@" -c $pkg/csrc/plan_inst.cu -o $pkg/build/ab/plan_h${H}_$name.o
objs=""
for o in $pkg/build/*.o; do
  case $o in */plan_h$H.o) objs="$objs $pkg/build/ab/plan_h${H}_$name.o";; *) objs="$objs $o";; esac
done
nvcc -shared -o $pkg/mbpo_b200/ab/libmbpo_$name.so $objs -cudart static
echo $pkg/mbpo_b200/ab/libmbpo_$name.so
