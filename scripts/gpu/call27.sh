set -x
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python scripts/time_kernels.py spmm --shapes mesh1,rgg --rowtile 4 --rtmodes 2,3 --reps 10 2>&1 | tee $O/spmm_modes_u8.txt
timeout 300 python bench.py --steps 30 --warmup 5 2>$O/bench_default.err | tail -1 > $O/bench_default.json
timeout 300 python bench.py --steps 2 --warmup 1 --no-graph --no-secondary --no-dp-check > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_mesh32k.csv python bench.py --steps 2 --warmup 1 --no-graph --no-secondary --no-dp-check > $O/launches_mesh32k.log 2>&1
cap() { # name regex skip args...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/ncu_$name python scripts/time_kernels.py "$@" --plain --reps 2 > $O/ncu_$name.log 2>&1
}
cap spmm_pipe_mesh1 spmm_step_rtile_pipe 6 spmm --shapes mesh1 --rowtile 4 --rtmodes 2
cap contract_fwd_tc3 contract_fwd_tc3 3 contract --shapes mesh1
cap contract_bwd_w_tc3 contract_bwd_w_tc3 3 contract --shapes mesh1
cap contract_bwd_x_tc3 contract_bwd_x_tc3 3 contract --shapes mesh2
cap bighead_bwd bighead_bwd 3 head
cap bighead_fc1 bighead_fc1 3 head
cap spmm_pipe_rgg spmm_step_rtile_pipe 6 spmm --shapes rgg --rowtile 4 --rtmodes 2
ls -la $O
