# round-2 profile pass: GPU tests, (8,16) and (3,3) bench lines, one ncu --set full capture of each step kernel
set -x
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
fi
timeout 300 python bench.py --steps 300 --warmup 20 --agents 8 --obstacles 16 --envs 262144 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>>gpurun_out/r2_bench.err > gpurun_out/r2_bench_8x16.json
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_8x16.json').read()); print('8x16 ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>>gpurun_out/r2_bench.err > gpurun_out/r2_bench_3x3.json
python -c "import json; d=json.loads(open('gpurun_out/r2_bench_3x3.json').read()); print('3x3 ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_team -s 25 -c 1 -o gpurun_out/r2_team_8x16 -f python bench.py --steps 10 --warmup 20 --agents 8 --obstacles 16 --envs 262144 --no-cpu-baseline --e2e-steps 1 --no-configs --no-strong > gpurun_out/ncu_team.log 2>&1; tail -3 gpurun_out/ncu_team.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_env -s 25 -c 1 -o gpurun_out/r2_env_3x3 -f python bench.py --steps 10 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-configs --no-strong > gpurun_out/ncu_env.log 2>&1; tail -3 gpurun_out/ncu_env.log
ls -la gpurun_out
