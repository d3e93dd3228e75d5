"""GPU: the fused inference path of the tower (cuDNN runtime-fused convolutions, merged head conv,
re-indexed dense GEMM) computes the same function as the plain PyTorch fp32 module of the same weights.
Tolerances: fused fp32 vs plain fp32 |dp| <= 2e-3 (TF32-free fp32 kernels differ only in summation
order); fused bf16 vs plain fp32 |dp| <= 0.15 (bf16 through 11 randomly initialised layers and a softmax)."""
import pytest
import torch

from galvanise_zero_b200 import bases, lookup
from galvanise_zero_b200.defs import datadesc
from galvanise_zero_b200 import nn as gznn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('game', ['reversi', 'breakthrough'])
def test_fused_tower_matches_fp32_reference(game):
    tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
    torch.manual_seed(0)
    net = gznn.make_inference_network(tr, 'small', seed=3, device='cuda', dtype=torch.bfloat16)
    ref = gznn.make_inference_network(tr, 'small', seed=3, device='cuda', dtype=torch.float32)
    ref.use_fused = False
    x = (torch.rand(257, tr.num_channels, tr.num_cols, tr.num_rows, device='cuda') > 0.7).float()
    f32 = gznn.make_inference_network(tr, 'small', seed=3, device='cuda', dtype=torch.float32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        p, v = net(x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))
        pf, vf = f32(x.contiguous(memory_format=torch.channels_last))
        pr, vr = ref(x)
    assert p.shape == pr.shape and v.shape == vr.shape
    assert torch.allclose(p.sum(1), torch.full((257,), float(len(tr.policy_dist_count)), device='cuda'), atol=1e-3)
    assert (pf - pr).abs().max().item() < 2e-3 and (vf - vr).abs().max().item() < 2e-3
    assert (p - pr).abs().max().item() < 0.15 and (v - vr).abs().max().item() < 0.15
