for v in "B2C_E2E_U8_CHUNK=8" "B2C_E2E_U8_CHUNK=12" "B2C_E2E_U8_CHUNK=6"; do
  echo "== $v"
  env $v python bench.py --steps 10 --cpu-budget 0.2 --no-train 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e_u8_input']['value'], d['e2e_u8_input']['ms_per_step'])"
done
