for v in "B2C_TRAIN_PRIORITY=1" "B2C_TRAIN_PRIORITY=0" "B2C_TRAIN_PRIORITY=1" "B2C_TRAIN_PRIORITY=0"; do
  echo "== $v"
  env $v python bench.py --steps 20 --cpu-budget 0.2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['train']['ms_per_step'], d['train']['value'])"
done
