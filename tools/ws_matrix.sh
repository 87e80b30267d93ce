# kernel-shape sweep over the bench configurations (one JSON line per shape and configuration)
SH="${SH:-8,8,1;8,12,1;10,8,1;8,8,2}"
run() { echo "## $*"; timeout 300 python tools/ws_sweep.py --steps 100 --reps 2 --shapes "$SH" "$@" 2>&1 | grep -o '"shape.*"ms_per_step": [0-9.]*'; }
run --experiment 8_arena --envs 65536
run --experiment 8_arena --envs 65536 --obs-dtype uint8
run --experiment 8_arena --envs 65536 --obs-dtype bfloat16
run --experiment 8_arena --envs 65536 --no-dense
run --experiment 8_arena --envs 16384
run --experiment 7_gridlocked --envs 16384
run --experiment 7_gridlocked --envs 65536
run --experiment 0_the_split --envs 65536
run --experiment 0_the_split --envs 4096
