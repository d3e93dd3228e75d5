python -m pytest tests/test_gpu_propnet.py -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2v_tests.log
rm -f gpurun_out/r2v_games.jsonl
for gs in "reversi 2097152" "connect6 262144" "amazons_10x10 131072" "chess_15d 65536" "hexLG13 8192" "breakthroughSmall 2097152" "breakthrough 2097152"; do
set -- $gs
python bench.py --workload propnet --game $1 --slots $2 --steps 300 --no-cpu-baseline >> gpurun_out/r2v_games.jsonl 2>> gpurun_out/r2v_games.err; echo "$1 rc=$?"
done
