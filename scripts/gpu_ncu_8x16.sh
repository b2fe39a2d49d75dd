set -x
mkdir -p gpurun_out
ARGS="--steps 20 --warmup 60 --no-cpu-baseline --e2e-steps 3 --agents 8 --obstacles 16 --envs 262144"
python bench.py $ARGS > gpurun_out/plain816.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_ -s 65 -c 2 -o gpurun_out/prof_step816 \
    python bench.py $ARGS > gpurun_out/ncu_full816.log 2>&1
tail -2 gpurun_out/ncu_full816.log
