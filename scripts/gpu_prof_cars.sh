mkdir -p gpurun_out
PROF="python scripts/gpu_cars_step.py"
$PROF > gpurun_out/cars_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_safe -s 8 -c 1 -f -o gpurun_out/prof_k_safe_cars $PROF > gpurun_out/ncu_cars.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_cars.log; cat gpurun_out/cars_plain.log | tail -1
