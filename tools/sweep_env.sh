# usage: bash tools/sweep_env.sh "VAR=a VAR2=b" "VAR=c" ...   (one bench run per argument)
for cfg in "$@"; do
  v=$(env $cfg timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['pcg']['iterations_per_step'])")
  echo "$cfg -> $v"
done
