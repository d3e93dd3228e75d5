python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 8 > gpurun_out/r2y_bench_compact.json 2> gpurun_out/r2y_bench.err; echo "bench compact rc=$?"; tail -3 gpurun_out/r2y_bench.err
python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 8 --full-batches > gpurun_out/r2y_bench_full.json 2>> gpurun_out/r2y_bench.err; echo "bench full rc=$?"
