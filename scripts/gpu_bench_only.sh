set -x
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python bench.py --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 3 2>>gpurun_out/bench_plain.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])"; done
