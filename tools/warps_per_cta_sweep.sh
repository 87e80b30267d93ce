# warps per CTA (libs built with -DCTF_WARPS_PER_CTA=w as tools/bin/libctf_w<w>.so) x resident CTAs per SM, 8_arena float32 B = 65536
run() { # lib caps...
  lib=$1; shift
  for c in "$@"; do
    printf "%s cap=%s " $lib $c
    CTF_B200_LIB=tools/bin/libctf_$lib.so timeout 100 python tools/ws_sweep.py --steps 150 --reps 2 --shapes kstep --k-step-ctas $c 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
  done
}
run w4 5 6 4
run w2 8 9 10 11 12 14
run w1 16 18 20 22 24
run w8 2 3
