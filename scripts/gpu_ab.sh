set -x
for rep in 1 2; do for v in B V; do
for cfg in "--agents 3 --obstacles 3 --envs 1048576 --steps 500"; do
MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so timeout 300 python bench.py $cfg --warmup 20 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['config']['num_agents'], 'ms_per_step', round(d['ms_per_step']*1000,2))"
done; done; done
