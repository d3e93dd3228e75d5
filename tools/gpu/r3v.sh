rm -f gpurun_out/r3v_games.jsonl
for mw in 8 6 5 4; do
for gs in "reversi 2097152" "amazons_10x10 131072"; do
set -- $gs
GZ_PN_MIN_WARPS=$mw timeout 300 python bench.py --workload propnet --game $1 --slots $2 --steps 200 --no-cpu-baseline >> gpurun_out/r3v_games.jsonl 2>> gpurun_out/r3v_games.err; echo "$1 mw=$mw rc=$?"
done
done
