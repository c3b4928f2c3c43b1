# A/B of prebuilt library variants (ab/lib_NAME.so, scripts/mk_variant.sh) on the fused SimulatedCars step.
mkdir -p gpurun_out
for r in 1 2; do
  for v in "$@"; do
    if [ "$v" = default ]; then unset RCBF_LIB_PATH; else export RCBF_LIB_PATH=$PWD/ab/lib_$v.so; fi
    python scripts/gpu_cars_step.py 2>&1 | tail -1 | cut -c1-60 | sed "s/^/$v /"
  done
done
unset RCBF_LIB_PATH
