mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --cpu-seconds 0 2>gpurun_out/bench_quick.err | tee gpurun_out/bench_quick.json | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('value %.4e ms/step %.4f e2e %.3e hbm_frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
    x=d['extra']; print({k:(v['value'] if isinstance(v,dict) and 'value' in v else v) for k,v in x.items() if k not in ('last_step_stats',)})
"; tail -3 gpurun_out/bench_quick.err
