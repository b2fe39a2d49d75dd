set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'stats', d['config']['episode_events_in_timed_region'])"
tail -2 gpurun_out/bench_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 300 --warmup 10 --scaling strong > gpurun_out/bench_n8_strong.json 2>> gpurun_out/bench_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_n8_strong.json').read().strip().splitlines()[-1])
print('strong N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'])"
