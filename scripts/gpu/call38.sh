set -x
O=gpurun_out/r02; mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 1 --no-graph --no-secondary --no-dp-check > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_mesh32k_final.csv python bench.py --steps 2 --warmup 1 --no-graph --no-secondary --no-dp-check > $O/launches_mesh32k_final.log 2>&1
timeout 400 python bench.py --workload rgg1m --steps 1 --warmup 1 --no-dp-check > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_rgg1m_final.csv python bench.py --workload rgg1m --steps 1 --warmup 1 --no-dp-check > $O/launches_rgg1m_final.log 2>&1
cap() { name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/ncu_$name python scripts/time_kernels.py "$@" --plain --reps 2 > $O/ncu_$name.log 2>&1; }
cap bighead_fc1_v2 bighead_fc1 3 head
cap contract_bwd_w_tc3_fused contract_bwd_w_tc3 3 contract --shapes mesh1
cap contract_fwd_tc3_ni2 contract_fwd_tc3 3 contract --shapes mesh1
ls -la $O | tail -12
