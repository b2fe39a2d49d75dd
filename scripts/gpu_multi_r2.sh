# round-2 multi-GPU evidence: N = $1 ranks (weak line with the strong-split record; host-copy probe)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 500 --warmup 20 --no-cpu-baseline --no-configs > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1500 gpurun_out/r2_bench_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/d2h_probe.py > gpurun_out/r2_d2h_probe_n$N.json 2> gpurun_out/r2_d2h_probe_n$N.err
cat gpurun_out/r2_d2h_probe_n$N.json
timeout 120 python scripts/d2h_probe.py > gpurun_out/r2_d2h_probe_n1.json 2>/dev/null; cat gpurun_out/r2_d2h_probe_n1.json
