N=${1:-2}
mkdir -p gpurun_out
timeout 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 profiles/ddp_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -8
echo "rc=$?"
for v in "B2C_AR_IN_GRAPH=1" "B2C_AR_IN_GRAPH=0"; do
B2C_AR_IN_GRAPH=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 2> gpurun_out/bench_n$N.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['train']['ms_per_step'], d['train']['value'], d['train']['allreduce'])"
done
