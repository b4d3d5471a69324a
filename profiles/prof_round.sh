mkdir -p gpurun_out
set -x
python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/b_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/ncu1.log 2>&1
python profiles/run_infer_once.py 8 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'tc|entropy|quant' -f -o gpurun_out/infer_full python profiles/run_infer_once.py 8 1 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
python profiles/run_train_once.py 2 || exit 1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python profiles/run_train_once.py 2 > gpurun_out/ncu3.log 2>&1
B2C_TRAIN_GRAPH=0 python bench.py --steps 5 > gpurun_out/bench_nograph.json 2>/dev/null
