python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in "B2C_E2E_CHUNK=4" "B2C_E2E_CHUNK=2" "B2C_E2E_CHUNK=3" "B2C_E2E_CHUNK=1"; do
  echo "== $v"
  env $v python bench.py --steps 10 --cpu-budget 0.2 --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"
done
