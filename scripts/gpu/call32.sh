set -x
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc_engine.py tests/test_gpu_properties.py tests/test_gpu_fullsize.py tests/test_gpu_model.py -x -q -m gpu 2>&1 | tail -3
for f in 1 0; do echo "FUSEA=$f"; TGCN_T3_FUSEA=$f timeout 300 python scripts/time_kernels.py contract --shapes mesh1,mesh2 --reps 10 2>&1 | grep "bwd_w"; done | tee $O/bwd_w_fused_a.txt
for f in 1 0; do echo "PDL=$f"; TGCN_SPMM_PDL=$f timeout 300 python scripts/time_kernels.py spmm --shapes mesh1,mesh2 --rowtile 4 --rtmodes 2 --reps 20 2>&1 | grep spmm; done | tee $O/spmm_pdl.txt
for f in 1 0; do TGCN_SPMM_PDL=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-secondary --no-dp-check 2>/dev/null | tail -1 > $O/bench_mesh32k_pdl$f.json; python -c "
import json;d=json.load(open('$O/bench_mesh32k_pdl$f.json'));print('PDL=$f', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['us_per_launch'], d['roofline']['frac'])"; done
