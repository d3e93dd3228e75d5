TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 500 $TR bench.py --gpus 8 --game hexLG13 --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3q_n8_hexLG13.json 2> gpurun_out/r3q_n8_hexLG13.err; echo "hex rc=$?"
timeout 500 $TR bench.py --gpus 8 --game chess_15d --advances-per-step 64 --presteps 300 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r3q_n8_chess_15d.json 2> gpurun_out/r3q_n8_chess_15d.err; echo "chess rc=$?"
