rm -f gpurun_out/r3g_games.jsonl
for gs in "reversi 2097152" "connect6 262144" "amazons_10x10 131072" "reversi 2097152" "connect6 262144"; do
set -- $gs
timeout 300 python bench.py --workload propnet --game $1 --slots $2 --steps 300 --no-cpu-baseline >> gpurun_out/r3g_games.jsonl 2>> gpurun_out/r3g_games.err; echo "$1 rc=$?"
done
