timeout 60 python -m pytest tests/test_encoder.py tests/test_training_batch.py -m gpu -x -q > gpurun_out/r5g_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5g_tests.log
timeout 40 python tools/bench_encoder.py > gpurun_out/r5g_encoder.jsonl 2> gpurun_out/r5g_encoder.err; echo "enc rc=$?"
timeout 45 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/r5g_bench.json 2> gpurun_out/r5g_bench.err; echo "bench rc=$?"
