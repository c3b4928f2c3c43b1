mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in 0 1 0 1; do echo "RCBF_NO_PDL=$v"; RCBF_NO_PDL=$v python bench.py --steps 30 --warmup 5 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    x=d['extra']
    print('  value %.4e ms/step %.4f  b512 fwd %.1f us fwdbwd %.1f us cars %.1f us  qp_uni %.3e cars_step %.3e'%(d['value'], d['ms_per_step'], x['config_unicycle_b512_fwd_us'], x['config3_unicycle_b512_fwd_bwd_us'], x['config2_cars_b512_fwd_us'], x['qp_solves_unicycle']['value'], x['cars_safe_step']['value']))
"; done
