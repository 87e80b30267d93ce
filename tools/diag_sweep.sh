# A/B of build variants of the extension (tools/bin/libctf_<name>.so) over kernel shapes; usage: bash tools/diag_sweep.sh "names" "shapes" [extra ws_sweep args]
for d in $1; do
  if [ $d = base ]; then unset CTF_B200_LIB; else export CTF_B200_LIB=tools/bin/libctf_$d.so; fi
  echo "== $d"
  timeout 200 python tools/ws_sweep.py --steps 100 --reps 2 --shapes "$2" $3 2>&1 | grep -o '"shape.*"ms_per_step": [0-9.]*'
done
