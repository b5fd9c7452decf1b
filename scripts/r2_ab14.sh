#!/bin/bash
# Round-2 A/B (gpurun): block sizes that give every SM the same number of cells in a single-wave launch; drop-in host path
mkdir -p gpurun_out
L=gpurun_out/r2_ab14.log; : > $L
for B in 128 136 113 116 128; do
  echo "== nat 100000 x 5yr block $B" >> $L
  HELP_B200_BLOCK=$B timeout 300 python scripts/dev_bench.py 100000 5 2>&1 | grep "^setup" | tail -1 >> $L
done
( timeout 600 python -m pytest tests/test_gpu_pack.py tests/test_abi.py tests/test_example_config.py -m gpu -x -q 2>&1 | tail -2 ) >> $L
timeout 900 python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
python - >> $L <<PY
import json
d=json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
print("bench value", d["value"], "e2e", d["e2e"]["value"], "dropin", d["e2e_dropin"]["value"], d["e2e_dropin"]["wall_s"])
PY
cat $L
