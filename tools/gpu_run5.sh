cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_superpoint_gpu.py tests/test_refpre_gpu.py tests/test_graph_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -n 6
