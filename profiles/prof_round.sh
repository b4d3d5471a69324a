# One profiling pass for profiles/ (run under gpurun, ONE GPU): bench line, ncu launch lists of the same commands,
# one --set full capture of the inference kernels, HBM-pass measurements.  Every ncu command runs only after the plain
# command exited 0; numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err || exit 1
python bench.py --steps 2 --warmup 3 --no-train --cpu-budget 0.2 > gpurun_out/b_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_infer.csv python bench.py --steps 2 --warmup 3 --no-train --cpu-budget 0.2 > gpurun_out/ncu1.log 2>&1
python profiles/run_infer_once.py 8 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'tc|entropy|quant' -f -o gpurun_out/infer_full python profiles/run_infer_once.py 8 1 > gpurun_out/ncu2.log 2>&1
python profiles/run_train_once.py 2 || exit 1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python profiles/run_train_once.py 2 > gpurun_out/ncu3.log 2>&1
python profiles/bench_passes.py > gpurun_out/passes.json
python profiles/train_layers_micro.py > gpurun_out/train_micro.txt 2>&1
ls -la gpurun_out
