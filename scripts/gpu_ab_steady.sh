# A/B of step-kernel builds with reset traffic established (400 warm-up steps): VARIANTS="PRE COOP" bash scripts/gpu_ab_steady.sh
mkdir -p gpurun_out
for rep in 1 2; do for v in ${VARIANTS:-PRE}; do
MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so timeout ${TMO:-120} python bench.py ${CFG:---agents 3 --obstacles 3 --envs 1048576} --steps 600 --warmup 400 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ABS $v', d['config']['num_agents'], 'us_per_step', round(d['ms_per_step']*1000,2))" | tee -a gpurun_out/ab.log
done; done
