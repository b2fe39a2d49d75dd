# occupancy sensitivity of the (8,16) team kernel: extra dynamic shared memory per CTA lowers the resident CTAs per SM
for pad in 0 512 1100 1800 2600 3600; do
MARLNAV_SMEM_PAD=$pad timeout 90 python bench.py --agents 8 --obstacles 16 --envs 262144 --steps 300 --warmup 20 --no-cpu-baseline --e2e-steps 3 --no-configs --no-strong 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['config']['smem_bytes']; print('OCC pad $pad smem', s, 'ctas/sm', 233472//(s+1024), 'us_per_step', round(d['ms_per_step']*1000,2))"
done
