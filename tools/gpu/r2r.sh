python tools/prof_puct.py 4096 3000 400 2>&1 | tail -1 | cut -c1-200
python tools/prof_puct.py 4096 3000 400 800 2>&1 | tail -1 | cut -c1-100
