#!/bin/bash
# scripts/ab_run.sh <repeats> <variant> [<variant> ...]: alternating bench runs of libraries built by ab_build.sh, on the
# three BASELINE.json single-GPU configs (step-kernel time only).
cd "$(dirname "$0")/.."
reps=$1; shift
for c in case33 case141 case322; do
  for r in $(seq $reps); do
    for v in "$@"; do
      MAPDN_B200_LIB=$PWD/scripts/variants/$v.so python bench.py --no-cpu --e2e-steps 3 --scenario $c 2>&1 | tail -1 |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c $v', round(d['ms_per_step']*1e3,2), 'us/step', round(d['value']/1e6,2), 'M env-steps/s')"
    done
  done
done
