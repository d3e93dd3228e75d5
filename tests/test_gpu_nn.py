"""GPU: the fused inference path of the tower (cuDNN runtime-fused convolutions, merged head conv,
re-indexed dense GEMM) computes the same function as the plain PyTorch fp32 module of the same weights.
Tolerances: merged heads vs plain heads on the same fp32 activations |dp| <= 5e-3; fused fp32 tower vs
plain fp32 tower <= 2 % of the activation range (cuDNN's runtime-fused fp32 kernels use TF32 tensor
cores); fused bf16 end to end vs plain fp32 |dp| <= 0.15 (bf16 through 11 randomly initialised layers
and a softmax)."""
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
        pr, vr = ref(x)
        # the merged head convolution + re-indexed dense GEMM against the plain heads, on the SAME
        # fp32 tower output (isolates the re-indexing from tensor-core rounding in the tower)
        xt = f32._tower(x.contiguous(memory_format=torch.channels_last))
        ph, vh = f32._heads_fused(xt)
        pp, vp = f32._heads_plain(xt)
        # the fused fp32 tower (cuDNN picks TF32 tensor-core kernels) against the plain fp32 tower
        xr = ref._tower(x)
    d_heads = max((ph - pp).abs().max().item(), (vh - vp).abs().max().item())
    d_tower = ((xt - xr).abs().max() / xr.abs().max()).item()
    d_bf16 = max((p - pr).abs().max().item(), (v - vr).abs().max().item())
    print('heads', d_heads, 'tower rel', d_tower, 'bf16 end-to-end', d_bf16)
    assert p.shape == pr.shape and v.shape == vr.shape
    assert torch.allclose(p.sum(1), torch.full((257,), float(len(tr.policy_dist_count)), device='cuda'), atol=1e-3)
    assert d_heads < 5e-3
    assert d_tower < 2e-2
    assert d_bf16 < 0.15
