#!/bin/bash
# Step-kernel time of the three BASELINE.json single-GPU configs (run after every kernel change: a change that is
# neutral for case33 / 8 lanes per env can cost the wide groups of case141 / case322 a lot, and vice versa).
cd "$(dirname "$0")/.."
for c in case33 case141 case322; do python scripts/kernel_time.py $c; done
