for cfg in "12 700" "3 700" "0.7 700" "0.7 200" "12 3000" "3 3000" "0 700"; do
  set -- $cfg
  v=$(FEMOL_GMG_CLUSTER_MB=$1 FEMOL_GMG_SOLO_NODES=$2 timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['pcg']['iterations_per_step'])")
  echo "cluster_mb=$1 solo_nodes=$2 -> $v"
done
