python -m pytest tests/test_gpu_model.py -x -q -k "host_pipeline" 2>&1 | tail -2
for v in "B2C_E2E_CHUNK=4" "B2C_E2E_CHUNK=3" "B2C_E2E_CHUNK=6" "B2C_E2E_CHUNK=2"; do
  echo "== $v"
  env $v python bench.py --steps 10 --cpu-budget 0.2 --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['call'])"
done
