TESTS=1 TMO=60 VARIANTS="W1 N" bash scripts/gpu_ab.sh
CFG="--agents 8 --obstacles 16 --envs 262144 --steps 300" TESTS=0 TMO=60 VARIANTS="W1 N" NCU=N bash scripts/gpu_ab.sh
