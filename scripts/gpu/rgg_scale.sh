# usage: bash scripts/gpu/rgg_scale.sh N   (strong scaling of the 1M-vertex layer over N GPUs, halo rows over NVLink peer memory)
set -x
N=$1; O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload rgg1m --steps 10 --warmup 3 2>$O/bench_rgg1m_n$N.err | tail -1 > $O/bench_rgg1m_n$N.json
python -c "
import json;d=json.load(open('$O/bench_rgg1m_n$N.json'));print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'], d.get('impl_detail'))"
tail -5 $O/bench_rgg1m_n$N.err
