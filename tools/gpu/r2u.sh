python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 --pipelines 2 > gpurun_out/r2u_bench_p2.json 2>> gpurun_out/r2u_bench.err; echo "bench p2 rc=$?"
python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 --pipelines 2 --max-work 8 > gpurun_out/r2u_bench_p2w8.json 2>> gpurun_out/r2u_bench.err; echo "bench p2 w8 rc=$?"
python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r2u_bench_p1.json 2>> gpurun_out/r2u_bench.err; echo "bench p1 rc=$?"
python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2u_tests.log
