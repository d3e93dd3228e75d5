for w in 6 8 12 16 24 40; do
python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 --max-work $w > gpurun_out/r2t_bench_w$w.json 2>> gpurun_out/r2t_bench.err; echo "bench w=$w rc=$?"
done
