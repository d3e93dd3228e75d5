rm -f gpurun_out/r3k_games.jsonl
for gc in 18 14 12 9 6; do
GZ_GROUP_CLUSTERS=$gc timeout 400 python bench.py --workload propnet --game chess_15d --slots 65536 --steps 100 --no-cpu-baseline >> gpurun_out/r3k_games.jsonl 2>> gpurun_out/r3k_games.err; echo "chess gc=$gc rc=$?"
done
for gc in 18 12 9; do
GZ_GROUP_CLUSTERS=$gc timeout 400 python bench.py --workload propnet --game hexLG13 --slots 18432 --steps 100 --no-cpu-baseline >> gpurun_out/r3k_games.jsonl 2>> gpurun_out/r3k_games.err; echo "hex gc=$gc rc=$?"
done
