for e in 1 2 3 4; do echo "EPB $e"; MAPDN_EPB=$e python scripts/lat_sweep.py case141 32 2>&1 | grep -E "B=1024|B=65536"; done
for e in 1 2 3; do echo "EPB $e"; MAPDN_EPB=$e python scripts/lat_sweep.py case322 64 2>&1 | grep -E "B=1024|B=16384"; done
