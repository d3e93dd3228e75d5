timeout 400 python -m pytest tests/test_training_batch.py tests/test_cabi.py tests/test_symmetry.py -m gpu -x -q > gpurun_out/r5a_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5a_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 5 > gpurun_out/r5a_bench.json 2> gpurun_out/r5a_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r5a_bench.err
