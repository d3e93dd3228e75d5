python -m pytest tests -m gpu -x -q > gpurun_out/r3j_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r3j_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r3j_smoke.log
python bench.py > gpurun_out/r3j_bench.json 2> gpurun_out/r3j_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3j_bench.err
python bench.py --impl reference > gpurun_out/r3j_ref.json 2> gpurun_out/r3j_ref.err; echo "ref rc=$?"
ncu --set full --clock-control none --import-source on -k regex:puct_pre_kernel -s 1550 -c 1 -o gpurun_out/r3j_pre -f python bench.py --steps 1 --warmup 3 --advances-per-step 8 --no-graph --no-e2e --no-cpu-baseline --no-propnet --presteps 1500 > gpurun_out/r3j_ncu_pre.log 2>&1; echo "ncu pre rc=$?"
