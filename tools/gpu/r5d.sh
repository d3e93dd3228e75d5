timeout 200 python -m pytest tests/test_encoder.py -m gpu -x -q -k "column" > gpurun_out/r5d_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r5d_tests.log
timeout 200 python tools/bench_encoder.py > gpurun_out/r5d_encoder.jsonl 2> gpurun_out/r5d_encoder.err; echo "enc rc=$?"; tail -2 gpurun_out/r5d_encoder.err
