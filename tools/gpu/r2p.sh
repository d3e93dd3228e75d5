python -m pytest tests/test_gpu_api.py tests/test_extension_module.py -m gpu -q > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2p_tests.log
for pf in 0 1 0 1; do
GZ_PUCT_PREFETCH=$pf python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r2p_bench_pf${pf}_$RANDOM.json 2>> gpurun_out/r2p_bench.err; echo "bench pf=$pf rc=$?"
done
