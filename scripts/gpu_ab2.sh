# A/B of step-kernel builds for both headline shapes: VARIANTS="BASE U8" bash scripts/gpu_ab2.sh
set -x
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
fi
for rep in 1 2; do for v in ${VARIANTS:-BASE U8}; do
for CFG in "--agents 8 --obstacles 16 --envs 262144 --steps 300" "--agents 3 --obstacles 3 --envs 1048576 --steps 500"; do
MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so timeout ${TMO:-90} python bench.py $CFG --warmup 20 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('AB $v', d['config']['num_agents'], 'us_per_step', round(d['ms_per_step']*1000,2))" | tee -a gpurun_out/ab.log
done; done; done
