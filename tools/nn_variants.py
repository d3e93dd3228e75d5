"""Times variants of the residual tower forward (4096 x 5x8x8) on the GPU."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch, torch.nn.functional as F
from galvanise_zero_b200 import bases, lookup
from galvanise_zero_b200.defs import datadesc
from galvanise_zero_b200.nn import make_inference_network
torch.backends.cudnn.benchmark = True
tr = bases.GdlBasesTransformer(lookup.by_name('reversi'), datadesc.generation_for('reversi'))
G = 4096
def timeit(fn, reps=20):
    for _ in range(5): fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for dt in (torch.bfloat16, torch.float16):
    net = make_inference_network(tr, 'small', seed=1, dtype=dt)
    x = torch.randn(G, 5, 8, 8, device='cuda').to(dt).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        print(dt, 'eager modules', timeit(lambda: net(x)))
        # fused cudnn ops
        def fused():
            y = torch.cudnn_convolution_relu(x, net.stem.weight, net.stem.bias, (1, 1), (1, 1), (1, 1), 1)
            for c0, c1 in net.blocks:
                z = torch.cudnn_convolution_relu(y, c0.weight, c0.bias, (1, 1), (1, 1), (1, 1), 1)
                y = torch.cudnn_convolution_add_relu(z, c1.weight, y, 1.0, c1.bias, (1, 1), (1, 1), (1, 1), 1)
            return y
        try:
            print(dt, 'fused cudnn tower only', timeit(fused))
        except Exception as e:
            print(dt, 'fused failed:', repr(e)[:200])
        def tower_only():
            y = F.relu(net.stem(x))
            for c0, c1 in net.blocks:
                z = F.relu(c0(y)); y = F.relu(c1(z) + y)
            return y
        print(dt, 'eager tower only', timeit(tower_only))
        def one_conv():
            return net.blocks[0][0](x64)
        x64 = torch.randn(G, 64, 8, 8, device='cuda').to(dt).contiguous(memory_format=torch.channels_last)
        print(dt, 'one 64->64 conv', timeit(one_conv))
        # GEMM-equivalent of one conv (M=G*64, N=64, K=576)
        a = torch.randn(G * 64, 576, device='cuda').to(dt); b = torch.randn(576, 64, device='cuda').to(dt)
        print(dt, 'equivalent GEMM', timeit(lambda: a @ b))
