# last-look confirmation: GPU tests, smoke, both bench arms (no profiler)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/confirm_pytest_gpu.log 2>&1; tail -3 gpurun_out/confirm_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/confirm_smoke.log 2>&1; tail -1 gpurun_out/confirm_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/confirm_bench_reference.json 2> gpurun_out/confirm_bench_reference.err
timeout 600 python bench.py > gpurun_out/confirm_bench.json 2> gpurun_out/confirm_bench.err; tail -c 300 gpurun_out/confirm_bench.err
cat gpurun_out/confirm_bench.json | cut -c1-600
