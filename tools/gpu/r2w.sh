python -m pytest tests/test_gpu_puct.py tests/test_gpu_api.py -m gpu -x -q -k "compacted or supervisor or worker or poll" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2w_tests.log
python bench.py --no-cpu-baseline --no-propnet --steps 8 > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2w_bench.err
