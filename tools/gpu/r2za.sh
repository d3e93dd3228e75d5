python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 8 > gpurun_out/r2za_bench_compact.json 2> gpurun_out/r2za_bench.err; echo "bench compact rc=$?"; tail -3 gpurun_out/r2za_bench.err
python -m pytest tests/test_gpu_puct.py -m gpu -x -q -k "compacted" 2>&1 | tail -3
