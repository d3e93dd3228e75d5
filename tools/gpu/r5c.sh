timeout 400 python -m pytest tests/test_encoder.py tests/test_training_batch.py tests/test_cabi.py -m gpu -x -q > gpurun_out/r5c_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r5c_tests.log
rm -f gpurun_out/r5c_games.jsonl
for g in reversi connect6 amazons_10x10 chess_15d breakthrough; do
  timeout 120 python bench.py --workload propnet --game $g --no-cpu-baseline --steps 100 >> gpurun_out/r5c_games.jsonl 2>> gpurun_out/r5c_games.err; echo "$g rc=$?"
done
