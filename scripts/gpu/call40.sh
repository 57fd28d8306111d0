O=gpurun_out/r02; mkdir -p $O
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 2>$O/final_bench_n2.err | tail -1 > $O/final_bench_n2.json
python -c "
import json;d=json.load(open('$O/final_bench_n2.json'));print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('dp_check'))"
tail -3 $O/final_bench_n2.err
