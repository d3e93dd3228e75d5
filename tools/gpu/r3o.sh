timeout 600 python -m pytest tests/test_gpu_puct.py -m gpu -x -q -k "compacted" 2>&1 | tail -3
timeout 500 python bench.py --game hexLG13 --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3n_puct_hexLG13.json 2> gpurun_out/r3n_puct_hexLG13.err; echo "hex rc=$?"; tail -2 gpurun_out/r3n_puct_hexLG13.err
