#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): bench lines of both arms, ncu launch lists, ncu --set full of the
# dominant kernels.  Outputs under gpurun_out/; summaries are copied into profiles/ by hand (tools/ncu_summary.py).
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py > $O/r02h_bench.json 2> $O/r02h_bench.err
python bench.py --impl reference > $O/r02h_bench_reference.json 2> $O/r02h_bench_reference.err
python - <<'PY'
import sys
sys.path.insert(0, ".")
import bench
bench.write_fasta("/tmp/msa_in.fa", bench.make_msa_reads(16, 50000, 0))
PY
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02h_launches_bench_pair100k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-suite > $O/ncu1.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02h_launches_msa16x50k.csv tsta_b200/_build/TSTA_msa -i /tmp/msa_in.fa -o /tmp/o.txt > $O/ncu2.log 2>&1
$NCU --set full --import-source on -k regex:poa_rows_kernel -s 12 -c 1 -o $O/r02h_poa_rows_16x50k -f tsta_b200/_build/TSTA_msa -i /tmp/msa_in.fa -o /tmp/o.txt > $O/ncu3.log 2>&1
$NCU --set full --import-source on -k regex:poa_walk -s 12 -c 1 -o $O/r02h_poa_walk_16x50k -f tsta_b200/_build/TSTA_msa -i /tmp/msa_in.fa -o /tmp/o.txt > $O/ncu4.log 2>&1
$NCU --set full --import-source on -k regex:psa_pair16 -s 4 -c 1 -o $O/r02h_psa_pair16_100k -f python tools/quick_pair.py > $O/ncu5.log 2>&1
$NCU --set full --import-source on -k regex:"psa_batch16_dual_kernel|psa_batch_trace_kernel" --launch-skip 2 --launch-count 2 -o $O/r02h_batch250_trace -f python tools/profile_batch.py 200000 250 trace 2 > $O/ncu6.log 2>&1
tail -n 2 $O/ncu3.log; tail -n 2 $O/ncu5.log; tail -n 2 $O/ncu6.log
tail -c 600 $O/r02h_bench.err
