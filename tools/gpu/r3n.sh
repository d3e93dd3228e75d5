for g in chess_15d hexLG13; do
timeout 500 python bench.py --game $g --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3n_puct_$g.json 2> gpurun_out/r3n_puct_$g.err; echo "$g rc=$?"; tail -2 gpurun_out/r3n_puct_$g.err
done
