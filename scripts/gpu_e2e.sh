for ch in 8 16 32; do echo "chunks=$ch"; RCBF_E2E_CHUNKS=$ch python bench.py --steps 10 --warmup 3 --no-extra --cpu-seconds 0 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('  e2e %.4e value %.4e'%(d['e2e']['value'], d['value']))
"; done
