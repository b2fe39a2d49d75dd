# instruction counts / issue utilisation of step-kernel builds: VARIANTS="BASE U2" CFG="--agents 8 ..." bash scripts/gpu_ncu_inst.sh
mkdir -p gpurun_out
CFG="${CFG:---agents 8 --obstacles 16 --envs 262144}"
for n in ${VARIANTS:-BASE}; do
  MARLNAV_B200_LIB=$PWD/build_ab/lib$n.so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active --clock-control none -k regex:step_ -s 30 -c 1 --csv --log-file gpurun_out/ncu_inst_$n.csv python bench.py $CFG --steps 20 --warmup 20 --no-cpu-baseline --e2e-steps 1 --no-configs --no-strong > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_inst_$n.csv')) if len(r)>10]
h=rows[0]; i=h.index('Metric Name'); v=h.index('Metric Value')
print('NCU $n', {r[i].split('.')[0][-34:]: r[v] for r in rows[1:]})
PY
done
