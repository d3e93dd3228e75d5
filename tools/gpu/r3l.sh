rm -f gpurun_out/r3l_games.jsonl
for gs in "chess_15d 65536" "hexLG13 18432"; do
set -- $gs
timeout 400 python bench.py --workload propnet --game $1 --slots $2 --steps 100 --no-cpu-baseline >> gpurun_out/r3l_games.jsonl 2>> gpurun_out/r3l_games.err; echo "$1 rc=$?"
done
timeout 900 python -m pytest tests/test_gpu_propnet.py -m gpu -x -q -k "group" 2>&1 | tail -2
