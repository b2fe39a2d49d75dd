# where does the thread-per-agent mapping of the team of 3 stop paying?  (MARLNAV_TEAM3_MAX_ENVS)
for B in 16384 32768 65536 131072; do for T in 0 1048576; do
MARLNAV_TEAM3_MAX_ENVS=$T timeout 90 python bench.py --agents 3 --obstacles 3 --envs $B --steps 2000 --warmup 50 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TEAM3 envs $B thread_per_agent', $T > 0, 'us_per_step', round(d['ms_per_step']*1000,2), d['config']['grid'])"
done; done
