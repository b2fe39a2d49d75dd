# final round-1 evidence: default bench line, reference arm, 8x16 line, launch list, ncu captures
set -x
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_final_reference.json 2>> gpurun_out/bench_final.err
timeout 300 python bench.py --steps 300 --warmup 20 --agents 8 --obstacles 16 --envs 262144 --cpu-budget 6 > gpurun_out/bench_final_8x16.json 2>> gpurun_out/bench_final.err
timeout 300 python scripts/config_sweep.py > gpurun_out/config_sweep.jsonl 2>> gpurun_out/bench_final.err
ARGS="--steps 40 --warmup 220 --no-cpu-baseline --e2e-steps 3"
python bench.py $ARGS > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_ -s 230 -c 2 -o gpurun_out/prof_step \
    python bench.py $ARGS > gpurun_out/ncu_full.log 2>&1
python bench.py $ARGS > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch.log 2>&1
ARGS8="--steps 20 --warmup 60 --no-cpu-baseline --e2e-steps 3 --agents 8 --obstacles 16 --envs 262144"
python bench.py $ARGS8 > gpurun_out/plain816.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:step_ -s 65 -c 2 -o gpurun_out/prof_step816 \
    python bench.py $ARGS8 > gpurun_out/ncu_full816.log 2>&1
tail -c 600 gpurun_out/bench_final.json
