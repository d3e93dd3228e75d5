import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
torch.backends.cudnn.benchmark = True
G = 4096
def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
dt = torch.bfloat16
b = torch.zeros(64, device='cuda', dtype=dt)
planes = torch.randn(G, 5, 8, 8, device='cuda').to(dt)
for cin in (5, 8, 16, 32, 64):
    w = torch.randn(64, cin, 3, 3, device='cuda').to(dt).contiguous(memory_format=torch.channels_last)
    x = torch.randn(G, cin, 8, 8, device='cuda').to(dt).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        t_conv = timeit(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1))
        def padcopy():
            x8 = torch.zeros((G, cin, 8, 8), device='cuda', dtype=dt).contiguous(memory_format=torch.channels_last)
            x8[:, :5] = planes
            return x8
        t_pad = timeit(padcopy) if cin >= 5 else 0
    print('cin', cin, 'conv us', round(t_conv, 1), 'pad+copy us', round(t_pad, 1))
