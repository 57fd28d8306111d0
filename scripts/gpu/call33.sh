O=gpurun_out/r02; mkdir -p $O
T=tests/test_gpu_model.py::test_mesh_model_with_large_head_and_fused_fc1_update_trains_like_the_port
echo "== default"; TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | tail -25
echo "== PDL=0"; TGCN_SPMM_PDL=0 TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | tail -8
echo "== FUSEA=0"; TGCN_T3_FUSEA=0 TGCN_TEST_VERBOSE=1 timeout 600 python -m pytest $T -x -q -m gpu -s 2>&1 | grep -v Warning | tail -8
