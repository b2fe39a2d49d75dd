# N-GPU bench lines (torchrun, NCCL): weak scaling, and strong scaling at N = 8
set -x
mkdir -p gpurun_out
N=${N:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 500 --warmup 20 --e2e-steps 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_n$N.err
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --gpus $N --steps 500 --warmup 20 --e2e-steps 5 --scaling strong > gpurun_out/bench_n${N}_strong.json 2>> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n${N}_strong.json | cut -c1-400
fi
