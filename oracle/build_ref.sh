#!/bin/bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference (bxskdh/TSTA) straight from its sources where they lie
# (default /root/reference, read-only) into oracle/_ref/<isa>/TSTA_{psa,psa_notrace,msa}.
# The reference's own Makefile is not used (it builds in-tree and ends with `clean`);
# this is the same gcc recipe it encodes (Makefile:1-53), once per ISA the reference's
# SIMD macro layer supports (psa/psa.c:18-108): sse4.2 (stock), avx2, avx512.
# No reference source is copied into this repository; oracle/_ref/ is git-ignored.
set -euo pipefail
REF="${1:-/root/reference}"
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
[ -f "$REF/psa/psa.c" ] || { echo "reference not found at $REF" >&2; exit 3; }
declare -A FLAGS=(
  [sse]="-msse4.2"
  [avx2]="-mavx2 -msse4.2"
  [avx512]="-mavx512f -mavx512bw -mavx512dq -mavx512vl -mavx2 -msse4.2"
)
for isa in sse avx2 avx512; do
  d="$OUT/$isa"; mkdir -p "$d/obj"
  CF="-O3 -g ${FLAGS[$isa]} -I$REF"
  gcc -O3 -g -c "$REF/pthreadpool/pthreadpool.c" -o "$d/obj/pthreadpool.o"
  gcc -O3 -g -c "$REF/seqio.c" -o "$d/obj/seqio.o"
  gcc $CF -c "$REF/psa/psa.c" -o "$d/obj/psa.o" -DTRACE
  gcc $CF -c "$REF/psa/psa.c" -o "$d/obj/psa-notrace.o"
  gcc $CF -c "$REF/msa/msa.c" -o "$d/obj/msa.o"
  gcc $CF -c "$REF/msa/c-t-simd.c" -o "$d/obj/c-t-simd.o"
  gcc -O3 -g -I$REF -c "$REF/msa/topo.c" -o "$d/obj/topo.o"
  gcc -O3 -g -I$REF -c "$REF/msa/result.c" -o "$d/obj/result.o"
  gcc -O3 -g -o "$d/TSTA_psa" "$d/obj/psa.o" "$d/obj/pthreadpool.o" "$d/obj/seqio.o" -lz -lm -lpthread
  gcc -O3 -g -o "$d/TSTA_psa_notrace" "$d/obj/psa-notrace.o" "$d/obj/pthreadpool.o" "$d/obj/seqio.o" -lz -lm -lpthread
  gcc -O3 -g -o "$d/TSTA_msa" "$d/obj/msa.o" "$d/obj/c-t-simd.o" "$d/obj/topo.o" "$d/obj/result.o" \
      "$d/obj/pthreadpool.o" "$d/obj/seqio.o" -lz -lm -lpthread
  rm -rf "$d/obj"
done
echo "built reference into $OUT"
