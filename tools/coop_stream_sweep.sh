#!/bin/bash
# k_step with cooperative streaming (-DCTF_COOP_STREAM=1: the CTA's warps write its consecutive env blocks together, one after
# the other) x warps per CTA x resident CTAs per SM; libs tools/bin/libctf_{base,coop2,coop4,coop8,coop16}.so
run() { # lib extra-args caps...
  lib=$1; shift; extra=$1; shift
  for c in "$@"; do
    printf "%s %s cap=%s " $lib "$extra" $c
    CTF_B200_LIB=tools/bin/libctf_$lib.so timeout 100 python tools/ws_sweep.py --steps 150 --reps 2 --shapes kstep --k-step-ctas $c $extra 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
  done
}
for rep in 1 2; do
run base "" 5
run coop4 "" 5 4 6 3
run coop8 "" 2 3
run coop16 "" 1
run coop2 "" 9 10 8
done
run base "--obs-dtype uint8" 0
run coop4 "--obs-dtype uint8" 0 5
run base "--experiment 7_gridlocked --envs 16384" -1
run coop4 "--experiment 7_gridlocked --envs 16384" -1 5 7
