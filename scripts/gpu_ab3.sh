bash scripts/gpu_ab.sh default s2w8b2
for c in 4 8 16 32; do RCBF_E2E_CHUNKS=$c python bench.py --steps 20 --warmup 5 --no-extra --cpu-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); v=d['extra']['e2e_variants']
print('chunks $c e2e full %.3e minimal %.3e devgp %.3e' % (d['e2e']['value'], v['minimal_outputs_host_gp_inputs']['value'], v['minimal_outputs_device_gp']['value']))"; done
