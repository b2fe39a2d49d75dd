# ncu: launch list + full capture of the step kernel in steady state (after resets have started)
set -x
mkdir -p gpurun_out
ARGS="--steps 40 --warmup 220 --no-cpu-baseline --e2e-steps 3 $EXTRA"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_ -s 230 -c 2 -o gpurun_out/prof_step \
    python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1
