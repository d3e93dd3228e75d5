python -m pytest tests/test_extension_module.py -m gpu -q -k player > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2q_tests.log
python tools/dbg_tail.py 3000 > gpurun_out/r2q_tail.txt 2>&1; echo "tail rc=$?"; cat gpurun_out/r2q_tail.txt
