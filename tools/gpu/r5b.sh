timeout 400 python -m pytest tests/test_puctplayer.py -m gpu -x -q > gpurun_out/r5b_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r5b_tests.log
timeout 300 python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 3 > gpurun_out/r5b_bench.json 2> gpurun_out/r5b_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r5b_bench.err
