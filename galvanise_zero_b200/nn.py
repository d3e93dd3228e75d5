"""Policy/value network: the reference's residual tower as a PyTorch module.

Architecture follows ``get_network_model`` (``src/ggpzero/nn/model.py:154-296``) for the non-"features"
configuration the size hints produce (``src/ggpzero/defs/templates.py:20-70``): 3x3 stem conv (no bias)
+ BN + ReLU, ``residual_layers`` v1 blocks (conv-BN-ReLU-conv-BN-add-ReLU, ``model.py:47-76``), one
policy head per role (1x1 conv to 2 channels + BN + ReLU, flatten, dense softmax, ``model.py:223-241``)
and the AG0-style value head (1x1 conv to 1 channel without BN + ReLU, flatten, dense hidden + ReLU,
dense softmax over 2 or 3 outcomes, ``model.py:278-291``). "small" = 64 filters, 5 residual layers,
value hidden 256 (``templates.py:42-45``).

The module consumes the encoder's device-resident planes ``[N, C, H, W]`` (bf16) and returns
``policy`` f32 ``[N, sum(policy sizes)]`` (role rows concatenated) and ``value`` f32 ``[N, heads]`` —
exactly the buffers ``PuctEngine.advance`` takes; nothing crosses to the host. Tensor cores are used
only here (cuDNN/cuBLAS through PyTorch). There are no trained weights in the reference
(``data/.MISSING_LARGE_BLOBS``); benchmarks use fixed-seed random weights.
"""
from __future__ import annotations

import attr
import torch
import torch.nn as nn
import torch.nn.functional as F


@attr.s
class NNModelConfig(object):
    """Same fields as ``src/ggpzero/defs/confs.py:126-151``."""
    role_count = attr.ib(default=2)
    input_rows = attr.ib(default=8)
    input_columns = attr.ib(default=8)
    input_channels = attr.ib(default=8)
    residual_layers = attr.ib(default=8)
    cnn_filter_size = attr.ib(default=64)
    cnn_kernel_size = attr.ib(default=3)
    value_hidden_size = attr.ib(default=256)
    policy_dist_count = attr.ib(default=attr.Factory(list))
    dropout_rate_policy = attr.ib(default=0.333)
    dropout_rate_value = attr.ib(default=0.5)
    leaky_relu = attr.ib(default=False)
    squeeze_excite_layers = attr.ib(default=False)
    resnet_v2 = attr.ib(default=False)
    global_pooling_value = attr.ib(default=False)
    concat_all_layers = attr.ib(default=False)


def nn_model_config_template(transformer, network_size_hint='small'):
    """``templates.nn_model_config_template`` (features=False)."""
    conf = NNModelConfig()
    conf.role_count = transformer.role_count
    conf.input_rows = transformer.num_rows
    conf.input_columns = transformer.num_cols
    conf.input_channels = transformer.num_channels
    conf.policy_dist_count = list(transformer.policy_dist_count)
    conf.cnn_kernel_size = 3
    conf.dropout_rate_policy = 0.25
    conf.dropout_rate_value = 0.5
    sizes = {'small': (64, 5, 256), 'medium': (96, 5, 512), 'large': (96, 10, 512)}
    conf.cnn_filter_size, conf.residual_layers, conf.value_hidden_size = sizes[network_size_hint]
    return conf


class ConvBN(nn.Module):
    def __init__(self, cin, cout, k, bn=True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=k // 2, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01) if bn else None   # Keras BN defaults

    def forward(self, x):
        x = self.conv(x)
        return self.bn(x) if self.bn is not None else x

    def fold(self):
        """Inference form: BN folded into a biased convolution."""
        conv = nn.Conv2d(self.conv.in_channels, self.conv.out_channels, self.conv.kernel_size,
                         padding=self.conv.padding, bias=True)
        w = self.conv.weight.detach().float()
        if self.bn is not None:
            scale = self.bn.weight.detach().float() / torch.sqrt(self.bn.running_var.float() + self.bn.eps)
            conv.weight.data = w * scale.view(-1, 1, 1, 1)
            conv.bias.data = self.bn.bias.detach().float() - self.bn.running_mean.float() * scale
        else:
            conv.weight.data = w
            conv.bias.data.zero_()
        return conv


class ResidualV1(nn.Module):
    def __init__(self, f, k):
        super().__init__()
        self.c0 = ConvBN(f, f, k)
        self.c1 = ConvBN(f, f, k)

    def forward(self, x):
        y = F.relu(self.c0(x))
        y = self.c1(y)
        return F.relu(x + y)


class ResidualTower(nn.Module):
    def __init__(self, conf: NNModelConfig, num_value_heads: int):
        super().__init__()
        assert not conf.resnet_v2 and not conf.squeeze_excite_layers and not conf.global_pooling_value, \
            'only the AG0-style (v1) tower of the size-hint templates is implemented'
        self.conf = conf
        f, k = conf.cnn_filter_size, conf.cnn_kernel_size
        self.stem = ConvBN(conf.input_channels, f, k)
        self.blocks = nn.ModuleList([ResidualV1(f, k) for _ in range(conf.residual_layers)])
        hw = conf.input_columns * conf.input_rows
        self.policy_convs = nn.ModuleList([ConvBN(f, 2, 1) for _ in conf.policy_dist_count])
        self.policy_dense = nn.ModuleList([nn.Linear(2 * hw, c) for c in conf.policy_dist_count])
        self.value_conv = ConvBN(f, 1, 1, bn=False)
        self.value_hidden = nn.Linear(hw, conf.value_hidden_size)
        self.value_out = nn.Linear(conf.value_hidden_size, num_value_heads)
        self.num_value_heads = num_value_heads

    def forward(self, planes):
        x = F.relu(self.stem(planes))
        for b in self.blocks:
            x = b(x)
        pols = []
        for conv, dense in zip(self.policy_convs, self.policy_dense):
            p = F.relu(conv(x)).flatten(1)
            pols.append(F.softmax(dense(p).float(), dim=1))
        v = F.relu(self.value_conv(x)).flatten(1)
        v = F.relu(self.value_hidden(v))
        v = F.softmax(self.value_out(v).float(), dim=1)
        return torch.cat(pols, dim=1).contiguous(), v.contiguous()

    def flops_per_position(self):
        c = self.conf
        hw = c.input_columns * c.input_rows
        f, k2 = c.cnn_filter_size, c.cnn_kernel_size ** 2
        fl = 2 * c.input_channels * f * k2 * hw + 2 * c.residual_layers * 2 * f * f * k2 * hw
        for cnt in c.policy_dist_count:
            fl += 2 * f * 2 * hw + 2 * 2 * hw * cnt
        fl += 2 * f * hw + 2 * hw * c.value_hidden_size + 2 * c.value_hidden_size * self.num_value_heads
        return fl


class InferenceTower(nn.Module):
    """BN-folded, bf16, channels-last inference copy of a ResidualTower."""

    def __init__(self, tower: ResidualTower):
        super().__init__()
        self.stem = tower.stem.fold()
        self.blocks = nn.ModuleList([nn.ModuleList([b.c0.fold(), b.c1.fold()]) for b in tower.blocks])
        self.policy_convs = nn.ModuleList([c.fold() for c in tower.policy_convs])
        self.policy_dense = tower.policy_dense
        self.value_conv = tower.value_conv.fold()
        self.value_hidden = tower.value_hidden
        self.value_out = tower.value_out
        self.use_fused = tower.conf.cnn_kernel_size == 3 and hasattr(torch, 'cudnn_convolution_add_relu')

    def _tower(self, planes):
        if planes.is_cuda and self.use_fused:
            # cuDNN runtime-fused conv+bias+ReLU and conv+bias+residual-add+ReLU: one kernel per
            # convolution instead of conv + 1-2 elementwise passes over the 33 MB activation
            one = (1, 1)
            cin = planes.shape[1]
            if cin != self.stem.in_channels:
                # the engine can hand over channels-last planes whose channel count is padded with zero
                # planes (PuctEngine(plane_c_pad=16)): pad the stem weights with zeros to match, so the
                # first convolution runs the same tensor-core kernel family as the tower
                if getattr(self, 'stem_wpad', None) is None or self.stem_wpad.shape[1] != cin:
                    w = torch.zeros((self.stem.out_channels, cin) + tuple(self.stem.kernel_size),
                                    device=planes.device, dtype=self.stem.weight.dtype)
                    w[:, :self.stem.in_channels] = self.stem.weight.detach()
                    self.stem_wpad = w.contiguous(memory_format=torch.channels_last)
                x = torch.cudnn_convolution_relu(planes, self.stem_wpad, self.stem.bias, one, one, one, 1)
            else:
                x = torch.cudnn_convolution_relu(planes, self.stem.weight, self.stem.bias, one, one, one, 1)
            for c0, c1 in self.blocks:
                y = torch.cudnn_convolution_relu(x, c0.weight, c0.bias, one, one, one, 1)
                x = torch.cudnn_convolution_add_relu(y, c1.weight, x, 1.0, c1.bias, one, one, one, 1)
            return x
        x = F.relu(self.stem(planes))
        for c0, c1 in self.blocks:
            y = F.relu(c0(x))
            x = F.relu(c1(y) + x)
        return x

    def _build_fused_heads(self):
        """One 1x1 convolution for all head channels (policy heads 2 each + value 1, padded to 8) and one
        GEMM for every dense layer that reads them, with the dense weights re-indexed to consume the
        channels-last activation directly (row = position*8 + channel)."""
        convs = list(self.policy_convs) + [self.value_conv]
        f = convs[0].in_channels
        n_ch = sum(c.out_channels for c in convs)
        assert n_ch <= 8
        dev, dt = convs[0].weight.device, convs[0].weight.dtype
        w = torch.zeros((8, f, 1, 1), device=dev, dtype=dt)
        b = torch.zeros((8,), device=dev, dtype=dt)
        ch0 = []
        c = 0
        for cv in convs:
            w[c:c + cv.out_channels] = cv.weight.detach()
            b[c:c + cv.out_channels] = cv.bias.detach()
            ch0.append(c)
            c += cv.out_channels
        self.head_w = w.contiguous(memory_format=torch.channels_last)
        self.head_b = b
        denses = list(self.policy_dense) + [self.value_hidden]
        hw = denses[0].in_features // convs[0].out_channels
        outs = [d.out_features for d in denses]
        big = torch.zeros((hw * 8, sum(outs)), device=dev, dtype=dt)
        bias = torch.zeros((sum(outs),), device=dev, dtype=dt)
        o = 0
        for d, cv, c0 in zip(denses, convs, ch0):
            wd = d.weight.detach()                      # [out, ch*hw], input index = ch_local*hw + pos
            for cl in range(cv.out_channels):
                rows = torch.arange(hw, device=dev) * 8 + (c0 + cl)
                big[rows, o:o + d.out_features] = wd[:, cl * hw:(cl + 1) * hw].t()
            bias[o:o + d.out_features] = d.bias.detach()
            o += d.out_features
        self.head_dense_w = big.contiguous()
        self.head_dense_b = bias
        self.head_splits = outs
        self.fused_heads = True

    def _heads_fused(self, x):
        if not getattr(self, 'fused_heads', False):
            self._build_fused_heads()
        one, zero = (1, 1), (0, 0)
        h = torch.cudnn_convolution_relu(x, self.head_w, self.head_b, one, zero, one, 1)   # [N,8,H,W] NHWC
        flat = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)                                # view: pos*8 + ch
        z = torch.addmm(self.head_dense_b, flat, self.head_dense_w)
        parts = torch.split(z, self.head_splits, dim=1)
        pols = [F.softmax(p.float(), dim=1) for p in parts[:-1]]
        v = F.relu(parts[-1])
        v = F.softmax(self.value_out(v).float(), dim=1)
        return torch.cat(pols, dim=1).contiguous(), v.contiguous()

    def _heads_plain(self, x):
        pols = []
        for conv, dense in zip(self.policy_convs, self.policy_dense):
            p = F.relu(conv(x)).flatten(1)
            pols.append(F.softmax(dense(p).float(), dim=1))
        v = F.relu(self.value_conv(x)).flatten(1)
        v = F.relu(self.value_hidden(v))
        v = F.softmax(self.value_out(v).float(), dim=1)
        return torch.cat(pols, dim=1).contiguous(), v.contiguous()

    def forward(self, planes):
        x = self._tower(planes)
        if planes.is_cuda and self.use_fused:
            return self._heads_fused(x)
        return self._heads_plain(x)


def make_inference_network(transformer, size='small', seed=0, device='cuda', dtype=torch.bfloat16):
    """Random-weight tower of the reference's architecture, ready for device-resident planes."""
    gen = torch.Generator().manual_seed(seed)
    conf = nn_model_config_template(transformer, size)
    tower = ResidualTower(conf, transformer.num_rewards)
    with torch.no_grad():
        for p in tower.parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=gen) * (2.0 / p[0].numel()) ** 0.5)
        for m in tower.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) + 0.5)
    tower.eval()
    net = InferenceTower(tower).to(device=device, dtype=dtype).to(memory_format=torch.channels_last)
    net.eval()
    net.flops_per_position = tower.flops_per_position()
    return net


class HeadResult(object):
    """``src/ggpzero/nn/network.py:9-17``."""

    def __init__(self, transformer, policies, values):
        assert len(transformer.policy_dist_count) == len(policies)
        self.policies = policies
        self.scores = values

    def __repr__(self):
        return "HeadResult(policies=%s, scores=%s" % (self.policies, self.scores)


class _HostBatchModel(object):
    """What the reference's pollers call: ``nn.get_model().predict_on_batch(planes)``
    (``src/ggpzero/util/cppinterface.py:119``) — host planes in, a list of host arrays
    ``[policy_0, .., policy_{R-1}, value]`` out. The forward is the PyTorch module on the device."""

    def __init__(self, module, transformer, device):
        self.module = module
        self.transformer = transformer
        self.device = torch.device(device)

    def predict_on_batch(self, planes):
        import numpy as np
        x = torch.from_numpy(np.ascontiguousarray(planes, dtype=np.float32))
        p = next(self.module.parameters())
        x = x.to(self.device, non_blocking=True).to(p.dtype)
        if self.device.type == 'cuda':
            x = x.contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            pol, val = self.module(x)
        pol = pol.float().cpu().numpy()
        val = val.float().cpu().numpy()
        out, off = [], 0
        for c in self.transformer.policy_dist_count:
            out.append(np.ascontiguousarray(pol[:, off:off + c]))
            off += c
        out.append(np.ascontiguousarray(val))
        return out

    predict = predict_on_batch


class NeuralNetwork(object):
    """The reference's ``NeuralNetwork`` surface on the hot path (``src/ggpzero/nn/network.py:20-68``):
    ``gdl_bases_transformer``, ``get_model()``, ``predict_n`` / ``predict_1``; the Keras model is
    replaced by the PyTorch tower. Training / compile / save are out of scope."""

    def __init__(self, gdl_bases_transformer, module, generation_descr=None, device='cuda'):
        self.gdl_bases_transformer = gdl_bases_transformer
        self.module = module
        self.generation_descr = generation_descr
        self._model = _HostBatchModel(module, gdl_bases_transformer, device)

    def get_model(self):
        return self._model

    def predict_n(self, states, prev_states=None):
        import numpy as np
        to_channels = self.gdl_bases_transformer.state_to_channels
        if prev_states:
            assert len(prev_states) == len(states)
            X = np.array([to_channels(s, prevs) for s, prevs in zip(states, prev_states)])
        else:
            X = np.array([to_channels(s) for s in states])
        Y = self._model.predict_on_batch(X)
        return [HeadResult(self.gdl_bases_transformer, [Y[k][i] for k in range(len(Y) - 1)], Y[-1][i])
                for i in range(len(states))]

    def predict_1(self, state, prev_states=None):
        if prev_states:
            return self.predict_n([state], [prev_states])[0]
        return self.predict_n([state])[0]
