# N-GPU bench line under torchrun (launched like the driver does)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-extra --cpu-seconds 0 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"; tail -2 gpurun_out/bench_n$N.err; python -c "
import json
for l in open('gpurun_out/bench_n$N.json'):
    try: d=json.loads(l)
    except Exception: continue
    print('N=%d value %.4e ms/step %.4f e2e %.3e stats %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['extra'].get('last_step_stats')))
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | tail -1 | cut -c1-300
