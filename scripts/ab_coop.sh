#!/bin/bash
# developer A/B (gpurun): the cooperative-lane mapping (HELP_B200_COOP) and the launch plan (HELP_B200_PLAN)
mkdir -p gpurun_out
L=gpurun_out/r2_coop_ab.log; : > $L
for G in 2 4 8; do echo "== parity COOP=$G" >> $L; HELP_B200_COOP=$G timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "families or rcra or ragged or six" 2>&1 | tail -2 >> $L; done
for G in 1 2 4 8; do echo "== example-size natural 17178 x 11yr COOP=$G" >> $L; HELP_B200_COOP=$G timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep simulate | tail -1 >> $L; done
echo "== example-size auto" >> $L; timeout 300 python scripts/dev_bench.py 17178 11 172 natural3 2>&1 | grep simulate | tail -1 >> $L
for G in 1 2 4; do echo "== landfill 31250 x 5yr COOP=$G" >> $L; HELP_B200_COOP=$G timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep simulate | tail -1 >> $L; done
for G in 1 2; do echo "== natural 50000 x 5yr COOP=$G" >> $L; HELP_B200_COOP=$G timeout 300 python scripts/dev_bench.py 50000 5 500 natural3 2>&1 | grep simulate | tail -1 >> $L; done
for P in 0 1; do for G in 1 2; do echo "== mixed 125000 x 5yr PLAN=$P COOP(wide)=$G" >> $L; HELP_B200_PLAN=$P HELP_B200_COOP=$G timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep simulate | tail -1 >> $L; done; done
echo "== mixed 125000 x 5yr PLAN=1 auto coop" >> $L; HELP_B200_PLAN=1 timeout 300 python scripts/dev_bench.py 125000 5 1250 mixed 2>&1 | grep simulate | tail -1 >> $L
echo "== bigreg A/B landfill" >> $L; HELP_B200_COOP=1 HELP_B200_LIB=$PWD/build_ab/lib_bigreg.so timeout 300 python scripts/dev_bench.py 31250 5 312 landfill 2>&1 | grep simulate | tail -1 >> $L
cat $L
