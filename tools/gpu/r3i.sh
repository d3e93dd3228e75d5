GZ_BENCH_CHILD_CAP=32768 python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r3i_bench_c32k.json 2>> gpurun_out/r3i_bench.err; echo "rc=$?"
GZ_BENCH_CHILD_CAP=32768 GZ_BENCH_NODE_CAP=3584 python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r3i_bench_c32k_n3584.json 2>> gpurun_out/r3i_bench.err; echo "rc=$?"
python bench.py --no-cpu-baseline --no-e2e --no-propnet --steps 8 > gpurun_out/r3i_bench_base.json 2>> gpurun_out/r3i_bench.err; echo "rc=$?"
