python tools/dbg_tail.py 40000 > gpurun_out/r2s_tail.txt 2>&1; echo "tail rc=$?"; head -24 gpurun_out/r2s_tail.txt; tail -6 gpurun_out/r2s_tail.txt
