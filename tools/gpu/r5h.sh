timeout 40 python -m pytest tests/test_encoder.py -m gpu -x -q -k "column" > gpurun_out/r5h_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5h_tests.log
GZ_BENCH_RAGGED=1 GZ_BENCH_SLOTS=262144 timeout 20 python tools/bench_encoder.py breakthroughSmall > gpurun_out/r5h_ragged.jsonl 2> gpurun_out/r5h_ragged.err; echo "bfs rc=$?"
