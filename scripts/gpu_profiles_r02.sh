# Round-2 evidence run (one box, one GPU): tests, bench (both arms), ncu launch list, full captures of the two hot kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
PROF="python bench.py --steps 20 --warmup 5 --no-extra --cpu-seconds 0"
$PROF > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_safe2 -s 6 -c 1 -f -o gpurun_out/prof_k_safe2 $PROF > gpurun_out/ncu_full2.log 2>&1; echo "ncu full rc=$?"
PROFB="python scripts/gpu_bwd_prof.py"
$PROFB > gpurun_out/bwd_plain.log 2>&1 && cat gpurun_out/bwd_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:bwd_tile -s 3 -c 1 -f -o gpurun_out/prof_bwd $PROFB > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
python scripts/gpu_latency.py > gpurun_out/latency.log 2>&1; tail -16 gpurun_out/latency.log
python scripts/gpu_cars_step.py > gpurun_out/cars_plain.log 2>&1 && cat gpurun_out/cars_plain.log | cut -c1-70 && \
ncu --set full --clock-control none --import-source on -k regex:k_cars2 -s 6 -c 1 -f -o gpurun_out/prof_k_cars2 python scripts/gpu_cars_step.py > gpurun_out/ncu_cars.log 2>&1; echo "ncu cars rc=$?"
python scripts/gpu_replay.py > gpurun_out/replay.log 2>&1; tail -10 gpurun_out/replay.log
