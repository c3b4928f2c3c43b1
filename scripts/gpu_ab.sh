# A/B of prebuilt library variants on one box: scripts/gpu_ab.sh NAME... (ab/lib_NAME.so; "default" = the in-tree build).
# Interleaved, 3 rounds each; prints ms_per_step of the headline kernel.
mkdir -p gpurun_out
for r in 1 2 3; do
  for v in "$@"; do
    if [ "$v" = default ]; then unset RCBF_LIB_PATH; else export RCBF_LIB_PATH=$PWD/ab/lib_$v.so; fi
    python bench.py --steps 20 --warmup 5 --no-extra --cpu-seconds 0 2> gpurun_out/ab_$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', r'$r', 'ms/step %.5f' % d['ms_per_step'], 'sustained %.5f' % d['sustained']['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'])"
  done
done
unset RCBF_LIB_PATH
