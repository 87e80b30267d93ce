for rep in 1 2; do for d in pad16 pad32 pad128; do
  export CTF_B200_LIB=tools/bin/libctf_$d.so
  for cfg in "--experiment 8_arena --envs 65536" "--experiment 7_gridlocked --envs 65536" "--experiment 7_gridlocked --envs 16384" "--experiment 8_arena --envs 16384"; do
  echo "== $d $cfg"
  timeout 200 python tools/ws_sweep.py --steps 100 --reps 2 --shapes "8,12,1" $cfg 2>&1 | grep -o '"shape.*"ms_per_step": [0-9.]*'
  done
done; done
