python -m pytest tests -m gpu -x -q > gpurun_out/r3u_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3u_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3u_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3u_smoke.log
python bench.py > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3u_bench.err
