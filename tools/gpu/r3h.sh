rm -f gpurun_out/r3h_games.jsonl
for gs in "hexLG13 18432" "hexLG13 36864" "chess_15d 147456"; do
set -- $gs
timeout 400 python bench.py --workload propnet --game $1 --slots $2 --steps 200 --no-cpu-baseline >> gpurun_out/r3h_games.jsonl 2>> gpurun_out/r3h_games.err; echo "$1 rc=$?"
done
tail -3 gpurun_out/r3h_games.err
