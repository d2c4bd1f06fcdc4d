cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_lightglue_gpu.py tests/test_parity_exact_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -n 8
bash tools/gpu_run6.sh
