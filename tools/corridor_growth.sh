#!/bin/bash
# traced 200 kbp / 400 kbp pairs of several seeds over corridor growth factors and margins
mkdir -p gpurun_out
{
for n in 200000 300000; do
for seed in 1 2 3 4; do
  for g in 200 141 100; do
    for c in "" 2048; do
      if [ -z "$c" ]; then unset TSTA_PSA_CORRIDOR; else export TSTA_PSA_CORRIDOR=$c; fi
      echo -n "n $n seed $seed growth $g margin ${c:-default}: "
      SEED=$seed TSTA_PSA_CORRIDOR_GROWTH=$g python tools/profile_one.py $n trace 3 2>&1 | tail -1
    done
  done
done
done
} > gpurun_out/corridor_growth.log 2>&1
