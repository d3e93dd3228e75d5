rm -f gpurun_out/r3f_games.jsonl
for gs in "reversi 2097152" "amazons_10x10 131072" "connect6 262144"; do
set -- $gs
timeout 300 python bench.py --workload propnet --game $1 --slots $2 --steps 300 --no-cpu-baseline >> gpurun_out/r3f_games.jsonl 2>> gpurun_out/r3f_games.err; echo "$1 rc=$?"
done
timeout 900 python -m pytest tests/test_gpu_propnet.py -m gpu -x -q 2>&1 | tail -2
