mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_closed_loop.py -m gpu -x -q -k "fused or cars or pending or host" 2>&1 | tail -5
for r in 1 2 3; do
  python scripts/gpu_cars_step.py 2>&1 | tail -1
  RCBF_NO_CARS2=1 python scripts/gpu_cars_step.py 2>&1 | tail -1 | sed 's/^/NO_CARS2 /'
done
