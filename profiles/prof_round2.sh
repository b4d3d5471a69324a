# Round-2 profiling pass for profiles/ (run under gpurun, ONE GPU): bench line, ncu launch list of the same command,
# one --set full capture of the inference kernels at the bench batch (24 images).  Every ncu command runs only after
# the plain command exited 0; numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err || exit 1
python bench.py --steps 2 --warmup 3 --no-train --cpu-budget 0.2 > gpurun_out/b_short.json 2>/dev/null || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_raw.csv python bench.py --steps 2 --warmup 3 --no-train --cpu-budget 0.2 > gpurun_out/ncu1.log 2>&1
python profiles/run_infer_once.py 24 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'tcp|conv1|entropy|quant' -s 9 -c 9 -f -o gpurun_out/r2_infer_full python profiles/run_infer_once.py 24 2 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out | tail -12
