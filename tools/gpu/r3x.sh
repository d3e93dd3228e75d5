timeout 900 python -m pytest tests/test_gpu_puct.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/r3x_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r3x_tests.log
timeout 500 python bench.py --game hexLG13 --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3x_puct_hexLG13.json 2> gpurun_out/r3x_puct_hexLG13.err; echo "hex rc=$?"; tail -2 gpurun_out/r3x_puct_hexLG13.err
timeout 500 python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r3x_bench.json 2> gpurun_out/r3x_bench.err; echo "reversi rc=$?"
