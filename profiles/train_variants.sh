mkdir -p gpurun_out
python -m pytest tests/test_gpu_model.py tests/test_gpu_tc.py -x -q 2>&1 | tail -3
for v in "B2C_TRAIN_KSPLIT=1" "B2C_TRAIN_KSPLIT=3" "B2C_TRAIN_KSPLIT=5" "B2C_TRAIN_KSPLIT=8"; do
  echo "== $v"
  env $v python bench.py --steps 10 --cpu-budget 0.5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['train']['ms_per_step'], d['train']['value'])"
done
