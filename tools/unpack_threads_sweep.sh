#!/bin/bash
# k_unpack: threads per CTA x chunk x residency, one box.  Needs tools/bin/libctf_u{128,256,512}.so (-DCTF_UNPACK_THREADS=...)
for rep in 1 2; do
for cfg in "u256 48 5" "u512 48 5" "u512 96 5" "u512 96 3" "u512 96 2" "u512 48 3" "u128 24 9" "u128 48 9" "u256 48 5"; do
  set -- $cfg
  echo "== lib=$1 chunk_kb=$2 ctas_per_sm=$3"
  CTF_B200_LIB=tools/bin/libctf_$1.so CTF_UNPACK_CHUNK_KB=$2 CTF_UNPACK_CTAS_PER_SM=$3 timeout 120 python tools/unpack_bench.py 2>&1 | grep -o 'dtype.*"GBps": [0-9.]*'
done; done
