python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
from galvanise_zero_b200.engine import PuctEngine
from galvanise_zero_b200.defs import confs
from galvanise_zero_b200.nn import make_inference_network
sp = confs.selfplay_config_template()
eng = PuctEngine('reversi', sp, n_games=4096, mode=0, node_cap=1024, child_cap=16384, plane_c_pad=16)
net = make_inference_network(eng.transformer, 'small', seed=1234, device=eng.device)
torch.backends.cudnn.benchmark = True
pc = eng.compact_requests(4096)
print('planes strides', eng.planes.stride(), 'planes_c strides', eng._planes_c.stride())
with torch.no_grad():
    for b in [2048, 2560, 3072, 3328, 3584, 3840, 4096]:
        x = eng._planes_c[:b]
        for _ in range(5): net(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            net(x)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): g.replay()
        e1.record(); torch.cuda.synchronize()
        print('rows', b, 'tower ms %.4f' % (e0.elapsed_time(e1)/50))
PY
python bench.py --no-cpu-baseline --no-propnet --no-e2e --steps 8 > gpurun_out/r2z_bench_compact.json 2> gpurun_out/r2z_bench.err; echo "bench compact rc=$?"; tail -3 gpurun_out/r2z_bench.err
