python bench.py --no-cpu-baseline --no-propnet --steps 8 --e2e-inline > gpurun_out/r2x_bench_inline.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2x_bench.err
