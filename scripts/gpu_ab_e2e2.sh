# A/B of the host-stepping (e2e) leg across libraries and chunk counts: VARIANTS="PRE TAIL" CHUNKS="8 16"
for rep in 1 2; do for v in ${VARIANTS:-PRE}; do for c in ${CHUNKS:-8}; do
MARLNAV_HOST_CHUNKS=$c MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so timeout 120 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 50 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('E2E $v chunks $c', round(d['e2e']['ms_per_step'],4), 'ms', round(d['e2e']['value']/1e6,1), 'M env-steps/s')"
done; done; done
