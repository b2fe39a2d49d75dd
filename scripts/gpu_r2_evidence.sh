# round-2 final evidence: GPU tests, smoke, default bench + reference arm, launch list, one ncu --set full capture per step kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
timeout 300 python bench.py --steps 300 --warmup 20 --agents 8 --obstacles 16 --envs 262144 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong > gpurun_out/r2_bench_8x16.json 2>>gpurun_out/r2_bench.err
timeout 300 python bench.py --steps 500 --warmup 20 --angle 3.14159265 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong > gpurun_out/r2_bench_trig_stress.json 2>>gpurun_out/r2_bench.err
# launch list of the default command's timed region (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_team -s 230 -c 1 -o gpurun_out/r2_final_team_8x16 -f python bench.py --steps 10 --warmup 225 --agents 8 --obstacles 16 --envs 262144 --no-cpu-baseline --e2e-steps 1 --no-configs --no-strong > gpurun_out/ncu_team.log 2>&1; tail -2 gpurun_out/ncu_team.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_env -s 230 -c 1 -o gpurun_out/r2_final_env_3x3 -f python bench.py --steps 10 --warmup 225 --no-cpu-baseline --e2e-steps 1 --no-configs --no-strong > gpurun_out/ncu_env.log 2>&1; tail -2 gpurun_out/ncu_env.log
ls -la gpurun_out | tail -20
