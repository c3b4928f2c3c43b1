# ncu launch list of one short bench run (one ncu invocation, after the same command ran clean)
mkdir -p gpurun_out
PROF="python bench.py --steps 3 --warmup 3 --no-extra --cpu-seconds 0"
$PROF > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|probe" --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"; tail -2 gpurun_out/ncu_launches.log
