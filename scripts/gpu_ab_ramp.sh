# A/B of the host leg's ramped first chunks (MARLNAV_HOST_RAMP) on build_ab/libRAMP.so
for rep in 1 2 3 4 5 6; do for r in 0 1; do
MARLNAV_HOST_RAMP=$r MARLNAV_B200_LIB=$PWD/build_ab/libRAMP.so timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 80 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('E2E ramp $r', round(d['e2e']['ms_per_step'],4), 'ms', round(d['e2e']['value']/1e6,1), 'M env-steps/s')"
done; done
