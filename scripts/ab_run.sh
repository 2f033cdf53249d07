#!/bin/bash
# scripts/ab_run.sh <repeats> <variant> [<variant> ...]: alternating bench runs of libraries built by ab_build.sh, on the
# three BASELINE.json single-GPU configs (step-kernel time only; `cur` = the in-tree library).
cd "$(dirname "$0")/.."
reps=$1; shift
for c in case33 case141 case322; do
  for r in $(seq $reps); do
    for v in "$@"; do
      lib=$PWD/scripts/variants/$v.so; [ "$v" = cur ] && lib=$PWD/mapdn_b200/libmapdn_b200.so
      MAPDN_B200_LIB=$lib python scripts/kernel_time.py $c | sed "s/^/$v /"
    done
  done
done
