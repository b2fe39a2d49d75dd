python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 3 --prewarm-s 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('FIRST prewarm0', 'us_per_step', round(d['ms_per_step']*1000,2))"
sleep 4
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 3 --prewarm-s 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('SECOND prewarm0', 'us_per_step', round(d['ms_per_step']*1000,2))"
sleep 4
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('THIRD prewarm0.5', 'us_per_step', round(d['ms_per_step']*1000,2))"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
