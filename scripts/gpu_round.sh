# tests + bench + ncu launch list + ncu full capture of the dominant kernel (1 GPU)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
PROF="python bench.py --steps 3 --warmup 3 --no-extra --cpu-seconds 0"
$PROF > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|probe" --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$PROF > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_safe -s 6 -c 2 -f -o gpurun_out/prof_k_safe $PROF > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
