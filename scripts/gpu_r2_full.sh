# round-2 full check: GPU tests, smoke, the default bench line, the reference arm as the driver runs it
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; tail -c 1200 gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_driver.json 2> gpurun_out/r2_bench_driver.err; tail -c 300 gpurun_out/r2_bench_driver.err
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_driver.json', 'gpurun_out/r2_bench.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'ms', round(d['ms_per_step'], 5), 'frac', round(d['roofline']['frac'], 4), 'fused_out', round(d['config']['step_fused_out_ms_per_step'], 5),
          'e2e', round(d['e2e']['value'] / 1e6, 1), 'M; cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'] / 1e6, 2), d['cpu_baseline']['kind']))
    print('  strong', d['strong'] and {k: d['strong'][k] for k in ('ms_per_step', 'value', 'eager_ms_per_step')})
    for k, v in (d['configs'] or {}).items():
        print('  ', k, {m: round(v[m]['ms_per_step'] * 1e3, 2) for m in ('step', 'step_fused_out', 'graph') if m in v}, 'us;', v.get('roofline', {}).get('frac'), v.get('rollout', {}).get('value'))
PY
