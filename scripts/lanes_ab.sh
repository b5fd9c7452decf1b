#!/bin/bash
# developer A/B: dense warps (HELP_B200_LANES=32) vs the batch spread evenly over the warps of a wave (default)
for n in 100000 94720 125000 75000; do
  for l in 32 auto; do
    if [ $l = auto ]; then unset HELP_B200_LANES; else export HELP_B200_LANES=$l; fi
    echo "== n=$n lanes=$l: $(timeout 300 python scripts/dev_bench.py $n 2 2>&1 | grep -E '^setup' | tail -1)"
  done
done
