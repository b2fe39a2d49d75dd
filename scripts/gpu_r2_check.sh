# round-2 iteration check: GPU parity tests, then the two headline benches (no CPU baseline)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 3 2>>gpurun_out/r2_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('3x3 ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
timeout 300 python bench.py --steps 300 --warmup 20 --agents 8 --obstacles 16 --envs 262144 --no-cpu-baseline --e2e-steps 3 2>>gpurun_out/r2_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('8x16 ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
