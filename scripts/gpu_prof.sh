# ncu --set full of the dominant kernel (one ncu invocation, after the same command ran clean)
mkdir -p gpurun_out
PROF="python bench.py --steps 3 --warmup 3 --no-extra --cpu-seconds 0"
$PROF > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_safe -s 6 -c 1 -f -o gpurun_out/prof_k_safe $PROF > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
