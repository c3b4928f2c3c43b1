# scripts/mk_variant.sh c2pend rcbf_cars2.cu -DRCBF_C2_FORCE_PENDING   (here, before the gpurun call)
mkdir -p gpurun_out
python scripts/gpu_cars_pending.py gpurun_out/cars_default.npz
RCBF_LIB_PATH=$PWD/ab/lib_c2pend.so python scripts/gpu_cars_pending.py gpurun_out/cars_pending.npz
python - <<'EOF'
import numpy as np
a, b = np.load("gpurun_out/cars_default.npz"), np.load("gpurun_out/cars_pending.npz")
bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
for k in bad:
    d = np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)); print(k, "differs in", int((d > 0).sum()), "entries, max", d.max())
print("forced-pending build == default build:", not bad)
EOF
