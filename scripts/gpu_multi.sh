# N-GPU bench (torchrun, NCCL) + reference arm
set -x
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
