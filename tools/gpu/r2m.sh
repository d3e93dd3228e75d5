python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2m_tests.log
python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
