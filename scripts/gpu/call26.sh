set -x
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q -m gpu -k "tiled or rowtile or variants or staged" 2>&1 | tail -5
timeout 300 python scripts/time_kernels.py spmm --shapes mesh1,mesh2 --rowtile 4 --rtmodes 1,2,3 --reps 20 2>&1 | tee gpurun_out/r02/spmm_modes.txt
timeout 300 python scripts/time_kernels.py spmm --shapes rgg --rowtile 4,8 --rtmodes 1,2 --reps 5 2>&1 | tee -a gpurun_out/r02/spmm_modes.txt
for m in 1 2 3; do TGCN_SPMM_RTILE=$m timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --no-dp-check 2>/dev/null | tail -1 > gpurun_out/r02/bench_mesh32k_rtile$m.json; python -c "
import json;d=json.load(open('gpurun_out/r02/bench_mesh32k_rtile$m.json'));print($m, d['ms_per_step'], d['roofline'])"; done
