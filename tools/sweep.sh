#!/bin/bash
# sweep strip width / chunk rows for the single-pair fill (experiments)
for tr in "" trace; do
for C in 4 8 16; do for CH in 8 16 32; do
  echo -n "C=$C CH=$CH $tr: "; TSTA_PSA_C=$C TSTA_PSA_CH=$CH python tools/profile_one.py ${1:-100000} $tr 3 | tail -n 1
done; done; done
