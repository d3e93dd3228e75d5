timeout 900 python -m pytest tests/test_gpu_propnet.py -m gpu -x -q > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3c_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-propnet --steps 8 > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3c_bench.err
