timeout 900 python -m pytest tests/test_gpu_api.py tests/test_extension_module.py -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r3d_tests.log
