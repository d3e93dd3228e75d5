timeout 200 python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 3 > gpurun_out/r5f_bench.json 2> gpurun_out/r5f_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r5f_bench.err
timeout 100 python tools/bench_encoder_one.py reversi > gpurun_out/r5f_one.log 2>&1; echo "one rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:encode_col_kernel -s 3 -c 1 -o gpurun_out/r5f_encode_col -f python tools/bench_encoder_one.py reversi > gpurun_out/r5f_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r5f_ncu.log
