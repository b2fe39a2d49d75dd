set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 200 --warmup 10 --cpu-budget 6 2>&1 | tail -5 | tee gpurun_out/bench_r1_first.json
