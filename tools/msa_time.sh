#!/bin/bash
# tools/msa_time.sh <nreads> <len>: phase timing of one TSTA_msa job on synthetic reads (twice; the second run is warm)
cd "$(dirname "$0")/.."
python - "$1" "$2" <<'PY'
import sys
sys.path.insert(0, ".")
import bench
reads = bench.make_msa_reads(int(sys.argv[1]), int(sys.argv[2]), 0)
bench.write_fasta("/tmp/msa_in.fa", reads)
PY
for rep in 1 2; do
  TSTA_MSA_TIMING=1 TSTA_POA_TIMING=1 tsta_b200/_build/TSTA_msa -i /tmp/msa_in.fa -o /tmp/msa_out.txt 2>&1 | grep -v "^\["
done
md5sum /tmp/msa_out.txt
