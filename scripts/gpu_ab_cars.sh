for v in "$@"; do echo "=== $v"; if [ "$v" = base ]; then L=$PWD/sac_rcbf_b200/librcbf_b200.so; else L=$PWD/sac_rcbf_b200/variants/librcbf_$v.so; fi
RCBF_LIB_PATH=$L python scripts/gpu_cars_step.py 2>&1 | tail -1 | cut -c1-70; done
