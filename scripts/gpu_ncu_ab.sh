# full ncu capture of the step kernel for each build in VARIANTS (libs in build_ab/lib<v>.so)
set -x
mkdir -p gpurun_out
ARGS="--steps 40 --warmup 220 --no-cpu-baseline --e2e-steps 3 ${CFG:-}"
for v in ${VARIANTS}; do
MARLNAV_B200_LIB=$PWD/build_ab/lib$v.so ncu --set full --clock-control none --import-source on -k regex:step_ -s 230 -c 1 -f -o gpurun_out/prof_$v \
    python bench.py $ARGS > gpurun_out/ncu_full_$v.log 2>&1
tail -2 gpurun_out/ncu_full_$v.log
done
