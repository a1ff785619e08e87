"""FlowNet2 (Ilg et al., CVPR 2017; NVIDIA's PyTorch port) around the B200 operators.

Same stack and state-dict keys as the reference's models/FlowNet/FlowNet2.py:18-177 with its
sub-networks FlowNetC.py, FlowNetS.py, FlowNetSD.py, FlowNetFusion.py in the configuration the
reference uses (batchNorm=False, fp16=False, rgb_max=255, div_flow=20; ownutilities.py:147-155).
The three custom operators — Correlation (FlowNetC.py:26-31), Resample2d x4 and ChannelNorm x6
(FlowNet2.py:126-167) — are this package's CUDA kernels; everything else is cuDNN/ATen.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
from torch.nn import init

_PAD = os.environ.get("PCFA_FN2_PAD", "1") != "0"


def conv(cin, cout, kernel_size=3, stride=1):
    from ..conv_ops import ConvLeakyReLU             # same children / state-dict keys as nn.Sequential(conv, LeakyReLU)
    return ConvLeakyReLU(nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=(kernel_size - 1) // 2, bias=True),
                         nn.LeakyReLU(0.1, inplace=True))


def i_conv(cin, cout, kernel_size=3, stride=1, bias=True):
    from ..conv_ops import ConvOnly
    return ConvOnly(nn.Conv2d(cin, cout, kernel_size=kernel_size, stride=stride, padding=(kernel_size - 1) // 2, bias=bias))


def predict_flow(cin):
    return nn.Conv2d(cin, 2, kernel_size=3, stride=1, padding=1, bias=True)


def deconv(cin, cout):
    from ..conv_ops import ConvLeakyReLU
    return ConvLeakyReLU(nn.ConvTranspose2d(cin, cout, kernel_size=4, stride=2, padding=1, bias=True),
                         nn.LeakyReLU(0.1, inplace=True))


def _cat(xs):
    """torch.cat(xs, 1).  On the GPU (channels-last fp32) one kernel that also appends zero channels up to a multiple of 8:
    the 1026-, 770-, 473-, 386-, 194-, 162-, 82- and 12-channel concatenations otherwise make cuDNN wrap every consuming
    convolution and its data gradient in channel-padding launches (202 per closure, 1.4 ms); consumers use zero-padded
    input-channel weights (conv_ops.apply_conv)."""
    x0 = xs[0]
    if x0.is_cuda and x0.dtype == torch.float32 and len(xs) <= 4 and _PAD:
        from ..gru_ops import cat_channels
        return cat_channels(list(xs), True, pad_to=8)
    return torch.cat(tuple(xs), 1)


def _skip(x):
    """(for the next encoder convolution, for the decoder's concatenation): conv_ops.fork on the GPU path."""
    from ..conv_ops import fork
    return fork(x) if _PAD else (x, x)


def _pf(m, x):
    """predict_flow / up-sampling (de)convolutions on a possibly zero-padded input."""
    from ..conv_ops import apply_conv
    return apply_conv(m, x, None) if x.is_cuda and _PAD else m(x)


def _xavier(mod):
    for m in mod.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            if m.bias is not None:
                init.uniform_(m.bias)
            init.xavier_uniform_(m.weight)


class _Refinement(nn.Module):
    """Decoder shared by FlowNetC / FlowNetS: deconv5..2, predict_flow6..2, upsampled_flow*."""

    def _make_decoder(self, up_bias):
        self.deconv5, self.deconv4, self.deconv3, self.deconv2 = deconv(1024, 512), deconv(1026, 256), deconv(770, 128), deconv(386, 64)
        for lvl, cin in ((6, 1024), (5, 1026), (4, 770), (3, 386), (2, 194)):
            setattr(self, f"predict_flow{lvl}", predict_flow(cin))
        for a, b in ((6, 5), (5, 4), (4, 3), (3, 2)):
            setattr(self, f"upsampled_flow{a}_to_{b}", nn.ConvTranspose2d(2, 2, 4, 2, 1, bias=up_bias))

    def _decode(self, c6, c5, c4, c3, c2):
        flow6 = _pf(self.predict_flow6, c6)
        cat5 = _cat((c5, self.deconv5(c6), _pf(self.upsampled_flow6_to_5, flow6)))
        flow5 = _pf(self.predict_flow5, cat5)
        cat4 = _cat((c4, self.deconv4(cat5), _pf(self.upsampled_flow5_to_4, flow5)))
        flow4 = _pf(self.predict_flow4, cat4)
        cat3 = _cat((c3, self.deconv3(cat4), _pf(self.upsampled_flow4_to_3, flow4)))
        flow3 = _pf(self.predict_flow3, cat3)
        cat2 = _cat((c2, self.deconv2(cat3), _pf(self.upsampled_flow3_to_2, flow3)))
        flow2 = _pf(self.predict_flow2, cat2)
        return (flow2, flow3, flow4, flow5, flow6) if self.training else (flow2,)


class FlowNetC(_Refinement):
    def __init__(self, correlation):
        super().__init__()
        self.conv1, self.conv2, self.conv3 = conv(3, 64, 7, 2), conv(64, 128, 5, 2), conv(128, 256, 5, 2)
        self.conv_redir = conv(256, 32, kernel_size=1, stride=1)
        self.corr = correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)
        self.corr_activation = nn.LeakyReLU(0.1, inplace=True)
        self.conv3_1 = conv(473, 256)
        self.conv4, self.conv4_1 = conv(256, 512, stride=2), conv(512, 512)
        self.conv5, self.conv5_1 = conv(512, 512, stride=2), conv(512, 512)
        self.conv6, self.conv6_1 = conv(512, 1024, stride=2), conv(1024, 1024)
        self._make_decoder(up_bias=True)
        _xavier(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')

    def forward(self, x):
        a2, a2s = _skip(self.conv2(self.conv1(x[:, 0:3])))
        a3 = self.conv3(a2)
        b3 = self.conv3(self.conv2(self.conv1(x[:, 3:])))
        corr = self.corr_activation(self.corr(a3, b3))
        c3, c3s = _skip(self.conv3_1(_cat((self.conv_redir(a3), corr))))
        c4, c4s = _skip(self.conv4_1(self.conv4(c3)))
        c5, c5s = _skip(self.conv5_1(self.conv5(c4)))
        c6 = self.conv6_1(self.conv6(c5))
        return self._decode(c6, c5s, c4s, c3s, a2s)


class FlowNetS(_Refinement):
    def __init__(self, input_channels=12):
        super().__init__()
        self.conv1, self.conv2, self.conv3 = conv(input_channels, 64, 7, 2), conv(64, 128, 5, 2), conv(128, 256, 5, 2)
        self.conv3_1 = conv(256, 256)
        self.conv4, self.conv4_1 = conv(256, 512, stride=2), conv(512, 512)
        self.conv5, self.conv5_1 = conv(512, 512, stride=2), conv(512, 512)
        self.conv6, self.conv6_1 = conv(512, 1024, stride=2), conv(1024, 1024)
        self._make_decoder(up_bias=False)
        _xavier(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')

    def forward(self, x):
        c2, c2s = _skip(self.conv2(self.conv1(x)))
        c3, c3s = _skip(self.conv3_1(self.conv3(c2)))
        c4, c4s = _skip(self.conv4_1(self.conv4(c3)))
        c5, c5s = _skip(self.conv5_1(self.conv5(c4)))
        c6 = self.conv6_1(self.conv6(c5))
        return self._decode(c6, c5s, c4s, c3s, c2s)


class FlowNetSD(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv0 = conv(6, 64)
        self.conv1, self.conv1_1 = conv(64, 64, stride=2), conv(64, 128)
        self.conv2, self.conv2_1 = conv(128, 128, stride=2), conv(128, 128)
        self.conv3, self.conv3_1 = conv(128, 256, stride=2), conv(256, 256)
        self.conv4, self.conv4_1 = conv(256, 512, stride=2), conv(512, 512)
        self.conv5, self.conv5_1 = conv(512, 512, stride=2), conv(512, 512)
        self.conv6, self.conv6_1 = conv(512, 1024, stride=2), conv(1024, 1024)
        self.deconv5, self.deconv4, self.deconv3, self.deconv2 = deconv(1024, 512), deconv(1026, 256), deconv(770, 128), deconv(386, 64)
        self.inter_conv5, self.inter_conv4 = i_conv(1026, 512), i_conv(770, 256)
        self.inter_conv3, self.inter_conv2 = i_conv(386, 128), i_conv(194, 64)
        for lvl, cin in ((6, 1024), (5, 512), (4, 256), (3, 128), (2, 64)):
            setattr(self, f"predict_flow{lvl}", predict_flow(cin))
        for a, b in ((6, 5), (5, 4), (4, 3), (3, 2)):
            setattr(self, f"upsampled_flow{a}_to_{b}", nn.ConvTranspose2d(2, 2, 4, 2, 1))
        _xavier(self)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')

    def forward(self, x):
        c1 = self.conv1_1(self.conv1(self.conv0(x)))
        c2, c2s = _skip(self.conv2_1(self.conv2(c1)))
        c3, c3s = _skip(self.conv3_1(self.conv3(c2)))
        c4, c4s = _skip(self.conv4_1(self.conv4(c3)))
        c5, c5s = _skip(self.conv5_1(self.conv5(c4)))
        c6 = self.conv6_1(self.conv6(c5))
        flow6 = _pf(self.predict_flow6, c6)
        cat5 = _cat((c5s, self.deconv5(c6), _pf(self.upsampled_flow6_to_5, flow6)))
        flow5 = _pf(self.predict_flow5, self.inter_conv5(cat5))
        cat4 = _cat((c4s, self.deconv4(cat5), _pf(self.upsampled_flow5_to_4, flow5)))
        flow4 = _pf(self.predict_flow4, self.inter_conv4(cat4))
        cat3 = _cat((c3s, self.deconv3(cat4), _pf(self.upsampled_flow4_to_3, flow4)))
        flow3 = _pf(self.predict_flow3, self.inter_conv3(cat3))
        cat2 = _cat((c2s, self.deconv2(cat3), _pf(self.upsampled_flow3_to_2, flow3)))
        flow2 = _pf(self.predict_flow2, self.inter_conv2(cat2))
        return (flow2, flow3, flow4, flow5, flow6) if self.training else (flow2,)


class FlowNetFusion(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv0 = conv(11, 64)
        self.conv1, self.conv1_1 = conv(64, 64, stride=2), conv(64, 128)
        self.conv2, self.conv2_1 = conv(128, 128, stride=2), conv(128, 128)
        self.deconv1, self.deconv0 = deconv(128, 32), deconv(162, 16)
        self.inter_conv1, self.inter_conv0 = i_conv(162, 32), i_conv(82, 16)
        self.predict_flow2, self.predict_flow1, self.predict_flow0 = predict_flow(128), predict_flow(32), predict_flow(16)
        self.upsampled_flow2_to_1 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        self.upsampled_flow1_to_0 = nn.ConvTranspose2d(2, 2, 4, 2, 1)
        _xavier(self)

    def forward(self, x):
        c0, c0s = _skip(self.conv0(x))
        c1, c1s = _skip(self.conv1_1(self.conv1(c0)))
        c2 = self.conv2_1(self.conv2(c1))
        flow2 = _pf(self.predict_flow2, c2)
        cat1 = _cat((c1s, self.deconv1(c2), _pf(self.upsampled_flow2_to_1, flow2)))
        flow1 = _pf(self.predict_flow1, self.inter_conv1(cat1))
        cat0 = _cat((c0s, self.deconv0(cat1), _pf(self.upsampled_flow1_to_0, flow1)))
        return _pf(self.predict_flow0, self.inter_conv0(cat0))


class FlowNet2(nn.Module):
    def __init__(self, ops=None, rgb_max=255.0, div_flow=20.):
        super().__init__()
        if ops is None:
            from .. import flownet2_ops as ops
        self.div_flow, self.rgb_max = div_flow, rgb_max
        self.channelnorm = ops.ChannelNorm()
        self.flownetc = FlowNetC(ops.Correlation)
        self.upsample1 = nn.Upsample(scale_factor=4, mode='bilinear')
        self.resample1 = ops.Resample2d()
        self.flownets_1 = FlowNetS()
        self.upsample2 = nn.Upsample(scale_factor=4, mode='bilinear')
        self.resample2 = ops.Resample2d()
        self.flownets_2 = FlowNetS()
        self.flownets_d = FlowNetSD()
        self.upsample3 = nn.Upsample(scale_factor=4, mode='nearest')
        self.upsample4 = nn.Upsample(scale_factor=4, mode='nearest')
        self.resample3 = ops.Resample2d()
        self.resample4 = ops.Resample2d()
        self.flownetfusion = FlowNetFusion()
        _xavier(self)

    def _refine_input(self, x, flow, resample):
        warped = resample(x[:, 3:], flow)
        err = self.channelnorm(x[:, :3] - warped)
        return _cat((x, warped, flow / self.div_flow, err))

    def forward(self, inputs):
        """inputs: [B, 3, 2, H, W] in [0, 255] (ownutilities.py:329-339)."""
        rgb_mean = inputs.contiguous().view(inputs.size()[:2] + (-1,)).mean(dim=-1).view(inputs.size()[:2] + (1, 1, 1))
        x = (inputs - rgb_mean) / self.rgb_max
        x = torch.cat((x[:, :, 0], x[:, :, 1]), dim=1)
        flow_c = self.upsample1(self.flownetc(x)[0] * self.div_flow)
        flow_s1 = self.upsample2(self.flownets_1(self._refine_input(x, flow_c, self.resample1))[0] * self.div_flow)
        flow_s2 = self.upsample4(self.flownets_2(self._refine_input(x, flow_s1, self.resample2))[0] * self.div_flow)
        norm_s2 = self.channelnorm(flow_s2)
        err_s2 = self.channelnorm(x[:, :3] - self.resample4(x[:, 3:], flow_s2))
        flow_sd = self.upsample3(self.flownets_d(x)[0] / self.div_flow)
        norm_sd = self.channelnorm(flow_sd)
        err_sd = self.channelnorm(x[:, :3] - self.resample3(x[:, 3:], flow_sd))
        return self.flownetfusion(torch.cat((x[:, :3], flow_sd, flow_s2, norm_sd, norm_s2, err_sd, err_s2), dim=1))
