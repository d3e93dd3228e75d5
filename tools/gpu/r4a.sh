timeout 400 python bench.py --no-cpu-baseline --no-propnet --steps 10 > gpurun_out/r4a_bench.json 2> gpurun_out/r4a_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r4a_bench.err
