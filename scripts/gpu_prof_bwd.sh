# ncu --set full of the compact backward kernel (after the same command ran clean)
mkdir -p gpurun_out
PROF="python scripts/gpu_bwd_prof.py"
$PROF > gpurun_out/bwd_plain.log 2>&1 && cat gpurun_out/bwd_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:bwd_meta -s 3 -c 1 -f -o gpurun_out/prof_bwd $PROF > gpurun_out/ncu_bwd.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_bwd.log
python scripts/gpu_bwd_prof.py 4194304 SimulatedCars
