T=tests/test_gpu_model.py::test_mesh_model_with_large_head_and_fused_fc1_update_trains_like_the_port
echo "== NI=4"; TGCN_T3_NI=4 TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | grep -E "worst|passed|failed" | head
echo "== NI=1"; TGCN_T3_NI=1 TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | grep -E "worst|passed|failed" | head
echo "== RTILE=1"; TGCN_SPMM_RTILE=1 TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | grep -E "worst|passed|failed" | head
echo "== V3 off"; TGCN_TC_V3=0 TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | grep -E "worst|passed|failed" | head
