set -x
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_head.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/time_kernels.py head --reps 20 2>&1 | tee $O/time_head.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --no-dp-check 2>/dev/null | tail -1 > $O/bench_mesh32k_b.json
python -c "
import json;d=json.load(open('$O/bench_mesh32k_b.json'));print(d['ms_per_step'], d['e2e'])"
timeout 600 python bench.py --workload rgg1m --steps 10 --warmup 3 2>$O/bench_rgg1m_n1.err | tail -1 > $O/bench_rgg1m_n1.json
python -c "
import json;d=json.load(open('$O/bench_rgg1m_n1.json'));print(d['ms_per_step'], d['e2e'], d['roofline'])"
