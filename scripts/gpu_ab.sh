# A/B of step-kernel builds on one box: VARIANTS="B V7" bash scripts/gpu_ab.sh  (libs in build_ab/lib<v>.so)
# TESTS=1 runs the GPU parity suite first (with the in-tree library); NCU=<variant> adds an instruction count.
set -x
mkdir -p gpurun_out
if [ "${TESTS:-0}" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
fi
CFG="${CFG:---agents 3 --obstacles 3 --envs 1048576 --steps 500}"
for rep in 1 2; do for v in ${VARIANTS:-B V7}; do
MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so timeout ${TMO:-90} python bench.py $CFG --warmup 20 --no-cpu-baseline --e2e-steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('AB $v', d['config']['num_agents'], 'us_per_step', round(d['ms_per_step']*1000,2))" | tee -a gpurun_out/ab.log
done; done
for n in ${NCU:-}; do
  MARLNAV_B200_LIB=$PWD/build_ab/lib$n.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:step_ -s 30 -c 1 --csv --log-file gpurun_out/ncu_inst_$n.csv python bench.py $CFG --steps 20 --warmup 20 --no-cpu-baseline --e2e-steps 3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_inst_$n.csv')) if len(r)>10]
h=rows[0]; i=h.index('Metric Name'); v=h.index('Metric Value')
print('NCU $n', {r[i].split('.')[0][-28:]: r[v] for r in rows[1:]})
PY
done
