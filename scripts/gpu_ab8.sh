# A/B of step-kernel builds at 262144 x 8 x 16 only: VARIANTS="NEW C26" bash scripts/gpu_ab8.sh
mkdir -p gpurun_out
for rep in 1 2; do for v in ${VARIANTS:-NEW}; do
MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so timeout ${TMO:-90} python bench.py ${CFG:---agents 8 --obstacles 16 --envs 262144 --steps 300} --warmup 20 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('AB $v', d['config']['num_agents'], 'us_per_step', round(d['ms_per_step']*1000,2))" | tee -a gpurun_out/ab.log
done; done
