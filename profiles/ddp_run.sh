# N-GPU checks of the gradient all-reduce; every run is bounded by `timeout`
N=${1:-2}
mkdir -p gpurun_out
timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/symm_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -12
echo "rc=$?"
for v in ${VARIANTS:-"B2C_AR=symm_graph" "B2C_AR=symm" "B2C_AR=nccl"}; do
  echo "== $v"
  env $v timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --cpu-budget 0.2 2> gpurun_out/bench_n$N.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['train']['ms_per_step'], d['train']['value'], d['train']['allreduce'][:60])"
  tail -3 gpurun_out/bench_n$N.err | cut -c1-300
done
