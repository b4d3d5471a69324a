# Round-2 (second session) profiling pass for profiles/ (run under gpurun, ONE GPU): bench line, ncu launch list of
# the same command, one --set full capture of the inference kernels at the bench batch (24 images), launch list of the
# graphed training step.  Every ncu command runs only after the plain command exited 0; numbers printed under ncu are
# never bench values.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err || exit 1
python bench.py --steps 2 --warmup 3 --no-train --cpu-budget 0.2 > gpurun_out/b_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_bench_raw.csv python bench.py --steps 2 --warmup 3 --no-train --cpu-budget 0.2 > gpurun_out/ncu1.log 2>&1
python profiles/run_infer_once.py 24 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'tcp|conv1|entropy|quant|deconv3x' -s 8 -c 10 -f -o gpurun_out/r2b_infer_full python profiles/run_infer_once.py 24 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
python profiles/run_train_once.py 2 || exit 1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r2b_launches_train_raw.csv python profiles/run_train_once.py 2 > gpurun_out/ncu3.log 2>&1
python profiles/trace_train.py gpurun_out/r2b_train_timeline.json > gpurun_out/r2b_train_timeline.txt 2>&1
ls -la gpurun_out | tail -12
