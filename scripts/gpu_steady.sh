# steady-state (reset traffic established) timing + instruction counts: warm-up of 400 steps first
mkdir -p gpurun_out
for CFG in "--agents 3 --obstacles 3 --envs 1048576" "--agents 8 --obstacles 16 --envs 262144"; do
for W in 20 400; do
timeout 120 python bench.py $CFG --steps 300 --warmup $W --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('STEADY warmup $W', d['config']['num_agents'], 'us_per_step', round(d['ms_per_step']*1000,2), d['config']['episode_events_in_timed_region'])"
done
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:step_ -s 430 -c 1 --csv --log-file gpurun_out/ncu_steady.csv python bench.py $CFG --steps 20 --warmup 420 --no-cpu-baseline --e2e-steps 1 --no-configs --no-strong > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_steady.csv')) if len(r)>10]
h=rows[0]; i=h.index('Metric Name'); v=h.index('Metric Value')
print('NCU steady $CFG', {r[i].split('.')[0][-34:]: r[v] for r in rows[1:]})
PY
done
