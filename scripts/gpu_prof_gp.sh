# ncu --set full of the three GP posterior kernels (one ncu invocation, after the same command ran clean)
mkdir -p gpurun_out
PROF="python scripts/gpu_gp.py"
$PROF > gpurun_out/gp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gp -c 12 -f -o gpurun_out/prof_k_gp $PROF > gpurun_out/ncu_gp.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_gp.log; tail -12 gpurun_out/gp_plain.log
