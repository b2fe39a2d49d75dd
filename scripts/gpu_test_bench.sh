# parity tests, then a plain bench line
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 500 --warmup 20 --cpu-budget 4 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err
cat gpurun_out/bench_plain.json; tail -3 gpurun_out/bench_plain.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --agents 8 --obstacles 16 --envs 262144 > gpurun_out/bench_8x16.json 2>> gpurun_out/bench_plain.err
cat gpurun_out/bench_8x16.json
