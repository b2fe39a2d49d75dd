# bench (plain) then ncu launch list + one full capture of the step kernel
set -x
mkdir -p gpurun_out
python bench.py --steps 500 --warmup 10 --cpu-budget 8 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err
cat gpurun_out/bench_plain.json; tail -3 gpurun_out/bench_plain.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 3 -o gpurun_out/prof_step \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
