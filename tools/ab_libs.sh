# usage: bash tools/ab_libs.sh "lib names (tools/bin/libctf_<name>.so)" "shapes" ["cfg1|cfg2|..."] -> ms per step
CFGS="${3:---experiment 8_arena --envs 65536|--experiment 8_arena --envs 65536 --obs-dtype uint8|--experiment 8_arena --envs 65536 --obs-dtype bfloat16|--experiment 7_gridlocked --envs 16384|--experiment 0_the_split --envs 65536|--experiment 0_the_split --envs 4096}"
for rep in 1 2; do for d in $1; do
  export CTF_B200_LIB=tools/bin/libctf_$d.so
  echo "$CFGS" | tr '|' '\n' | while read cfg; do
  echo "== $d $cfg"
  timeout 200 python tools/ws_sweep.py --steps 100 --reps 2 --shapes "$2" $cfg 2>&1 | grep -o '"shape.*"ms_per_step": [0-9.]*'
  done
done; done
