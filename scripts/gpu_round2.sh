# tests + quick bench + full ncu capture of the dominant kernel (k_safe2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
bash scripts/gpu_ab.sh base
PROF="python bench.py --steps 3 --warmup 3 --no-extra --cpu-seconds 0"
ncu --set full --clock-control none --import-source on -k regex:k_safe2 -s 6 -c 1 -f -o gpurun_out/prof_k_safe2 $PROF > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full2.log
