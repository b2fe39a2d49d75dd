TESTS=1 TMO=60 VARIANTS="N3 N4" NCU=N4 bash scripts/gpu_ab.sh
CFG="--agents 8 --obstacles 16 --envs 262144 --steps 300" TESTS=0 TMO=60 VARIANTS="N3 N4" bash scripts/gpu_ab.sh
