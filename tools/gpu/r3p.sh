python tools/dbg_tail.py 1200 hexLG13 1024 393216 > gpurun_out/r3p_tail_hex.txt 2>&1; echo "rc=$?"; head -30 gpurun_out/r3p_tail_hex.txt; tail -6 gpurun_out/r3p_tail_hex.txt
