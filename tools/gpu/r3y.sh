timeout 900 python -m pytest tests/test_gpu_propnet.py -m gpu -x -q -k "group" > gpurun_out/r3y_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3y_tests.log
timeout 600 python -m pytest tests/test_gpu_puct.py -m gpu -x -q -k "hexLG13 or chess" 2>&1 | tail -2
timeout 500 python bench.py --game hexLG13 --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3y_puct_hexLG13.json 2> gpurun_out/r3y_puct_hexLG13.err; echo "hex rc=$?"; tail -2 gpurun_out/r3y_puct_hexLG13.err
timeout 500 python bench.py --game chess_15d --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3y_puct_chess_15d.json 2> gpurun_out/r3y_puct_chess_15d.err; echo "chess rc=$?"
