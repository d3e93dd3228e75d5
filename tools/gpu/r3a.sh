timeout 900 python -m pytest tests/test_gpu_propnet.py tests/test_gpu_puct.py -m gpu -x -q > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3a_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 8 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3a_bench.err
timeout 300 python bench.py --workload propnet --game reversi --slots 2097152 --steps 300 --no-cpu-baseline > gpurun_out/r3a_pn_reversi.json 2>> gpurun_out/r3a_bench.err; echo "pn rc=$?"
