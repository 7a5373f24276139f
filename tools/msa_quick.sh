#!/bin/bash
# tools/msa_quick.sh <nreads> <len>: one TSTA_msa job on synthetic reads with kernel / phase timing, both fill kernels
cd "$(dirname "$0")/.."
python - "$1" "$2" <<'PY'
import sys
sys.path.insert(0, ".")
import bench
reads = bench.make_msa_reads(int(sys.argv[1]), int(sys.argv[2]), 0)
bench.write_fasta("/tmp/msa_in.fa", reads)
PY
for k in rows skew; do
  for rep in 1 2; do
    TSTA_POA_KERNEL=$k TSTA_MSA_TIMING=1 TSTA_POA_TIMING=1 tsta_b200/_build/TSTA_msa -i /tmp/msa_in.fa -o /tmp/msa_$k.txt 2>&1 | grep -v "^\[" | tr '\n' ' '; echo
  done
done
md5sum /tmp/msa_rows.txt /tmp/msa_skew.txt
