# chunk count of the host-stepping pipeline (marlnav_step_host_f32) at 1M envs
for rep in 1 2; do for c in 1 2 4 8 16; do
MARLNAV_HOST_CHUNKS=$c timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 50 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('E2E chunks $c', round(d['e2e']['ms_per_step'],4), 'ms', round(d['e2e']['value']/1e6,1), 'M env-steps/s')"
done; done
