#!/bin/bash
# A/B builds of libptb.so for GPU experiments: tools/build_variants.sh name1 "defs1" name2 "defs2" ...
# defs apply to every translation unit; a part after a '|' applies to the fast-arithmetic build only ("defs|fast defs")
# -> build/var_<name>/libptb.so (selected at run time with PTB_LIB=...)
cd "$(dirname "$0")/../szakdolgozat_pathtracer_b200/csrc"
pids=()
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  mkdir -p ../../build/var_$name
  all="${defs%%|*}"; fast=""; [[ "$defs" == *"|"* ]] && fast="${defs#*|}"
  ( make -s OBJDIR=../../build/var_$name/obj OUT=../../build/var_$name/libptb.so CLI=../../build/var_$name/ptb_render EXTRA_DEFS="$all" FAST_DEFS="$fast" ../../build/var_$name/libptb.so > ../../build/var_$name/build.log 2>&1 && echo "built $name" || echo "FAILED $name" ) &
  pids+=($!)
done
wait
