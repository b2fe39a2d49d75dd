# the driver's --steps 20 --warmup 5 line with the clock sampler armed inside the launch loop (legacy) vs after it
for rep in 1 2 3 4 5 6; do for m in legacy parked; do
MARLNAV_BENCH_SAMPLER=$m timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('SAMPLER $m', round(d['ms_per_step']*1e3,2), 'us', round(d['value']/1e9,2), 'G', d['clocks'])"
done; done
