# A/B of kernel build variants: one quick bench line per variant
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== variant $v"
  RCBF_LIB_PATH=$PWD/sac_rcbf_b200/variants/librcbf_$v.so python bench.py --steps 20 --warmup 5 --no-extra --cpu-seconds 0 2>gpurun_out/ab_$v.err | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  value %.4e ms/step %.4f e2e %.3e hbm_frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac']))
"
done
