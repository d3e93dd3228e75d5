python -m pytest tests -m gpu -x -q > gpurun_out/r3z_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r3z_tests.log
