for v in "$@"; do echo "=== $v"; if [ "$v" = base ]; then L=$PWD/sac_rcbf_b200/librcbf_b200.so; else L=$PWD/sac_rcbf_b200/variants/librcbf_$v.so; fi
RCBF_LIB_PATH=$L python bench.py --steps 20 --warmup 5 --no-extra --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  value %.4e ms/step %.4f'%(d['value'], d['ms_per_step']))
"; done
