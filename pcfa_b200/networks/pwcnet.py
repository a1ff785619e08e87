"""PWC-Net (PWC-DC variant; Sun et al., CVPR 2018) around the B200 cost-volume and warp operators.

Same architecture and state-dict keys as the reference's models/PWCNet/PWCNet.py:60-330.  The five
`correlate` calls (PWCNet.py:45-58; 9x9 sampler, /C) and the four `warp` calls (PWCNet.py:166-206)
run as fused CUDA kernels on the device the features live on — the reference's default moves every
correlation to the CPU and back (correlationSamplerOnlyCPU, PWCNet.py:18-21).  LeakyReLU after the
correlation and the decoder stay in cuDNN/ATen.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def conv(cin, cout, kernel_size=3, stride=1, padding=1, dilation=1):
    from ..conv_ops import ConvLeakyReLU             # same children / state-dict keys as nn.Sequential(conv, LeakyReLU)
    return ConvLeakyReLU(nn.Conv2d(int(cin), int(cout), kernel_size=kernel_size, stride=stride, padding=padding,
                                   dilation=dilation, bias=True), nn.LeakyReLU(0.1))


def predict_flow(cin):
    return nn.Conv2d(int(cin), 2, kernel_size=3, stride=1, padding=1, bias=True)


def deconv(cin, cout, kernel_size=4, stride=2, padding=1):
    return nn.ConvTranspose2d(int(cin), int(cout), kernel_size, stride, padding, bias=True)


_PYRAMID = [(1, 3, 16), (2, 16, 32), (3, 32, 64), (4, 64, 96), (5, 96, 128), (6, 128, 196)]
_FLOW_SCALE = {5: 0.625, 4: 1.25, 3: 2.5, 2: 5.0}          # PWCNet.py:263,277,291,307


class PWCDCNet(nn.Module):
    def __init__(self, md=4, ops=None):
        super().__init__()
        if ops is None:
            from ..pwc_warp import pwc_warp
            from ..spatial_correlation_sampler import pwc_correlate
            self.corr, self.warp = pwc_correlate, pwc_warp
        else:
            self.corr, self.warp = ops.pwc_correlate, ops.pwc_warp
        self.upsample = nn.Upsample(scale_factor=4, mode='bilinear')
        for lvl, cin, cout in _PYRAMID:
            first, second = ("aa", "a") if lvl == 6 else ("a", "aa")     # level 6 is named conv6aa, conv6a, conv6b
            setattr(self, f"conv{lvl}{first}", conv(cin, cout, stride=2))
            setattr(self, f"conv{lvl}{second}", conv(cout, cout))
            setattr(self, f"conv{lvl}b", conv(cout, cout))
        self.leakyRELU = nn.LeakyReLU(0.1)
        nd = (2 * md + 1) ** 2
        dd = np.cumsum([128, 128, 96, 64, 32])
        feat = {6: 0, 5: 128, 4: 96, 3: 64, 2: 32}
        for lvl in (6, 5, 4, 3, 2):
            od = nd + (feat[lvl] + 4 if lvl < 6 else 0)
            for i, (extra, width) in enumerate(zip([0, *dd[:4]], [128, 128, 96, 64, 32])):
                setattr(self, f"conv{lvl}_{i}", conv(od + extra, width))
            setattr(self, f"predict_flow{lvl}", predict_flow(od + dd[4]))
            setattr(self, f"deconv{lvl}", deconv(2, 2))
            if lvl > 2:
                setattr(self, f"upfeat{lvl}", deconv(od + dd[4], 2))
        od = nd + 32 + 4
        self.dc_conv1 = conv(od + dd[4], 128, padding=1, dilation=1)
        self.dc_conv2 = conv(128, 128, padding=2, dilation=2)
        self.dc_conv3 = conv(128, 128, padding=4, dilation=4)
        self.dc_conv4 = conv(128, 96, padding=8, dilation=8)
        self.dc_conv5 = conv(96, 64, padding=16, dilation=16)
        self.dc_conv6 = conv(64, 32, padding=1, dilation=1)
        self.dc_conv7 = predict_flow(32)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight.data, mode='fan_in')
                if m.bias is not None:
                    m.bias.data.zero_()

    def _features(self, im):
        feats = {}
        x = im
        for lvl, _, _ in _PYRAMID:
            first, second = ("aa", "a") if lvl == 6 else ("a", "aa")
            x = getattr(self, f"conv{lvl}b")(getattr(self, f"conv{lvl}{second}")(getattr(self, f"conv{lvl}{first}")(x)))
            feats[lvl] = x
        return feats

    def _cat(self, xs, pad=False):
        """torch.cat(xs, 1); on the channels-last path one vectorised kernel (ATen's channels-last cat is a slow path).
        pad=True appends zero channels up to a multiple of 8: the decoder's 213-, 181-, 149-, 117-channel inputs (and every
        DenseNet concatenation grown from them) otherwise make cuDNN wrap each convolution and data gradient in
        channel-padding launches (150 per closure, 0.9 ms); the consumers use zero-padded input-channel weights."""
        if self._cl(xs[0]):
            from ..gru_ops import cat_channels
            return cat_channels(list(xs), True, pad_to=8 if pad else 0)
        return torch.cat(tuple(xs), 1)

    def _plain(self, m, x):
        """A plain Conv2d / ConvTranspose2d of the decoder on a possibly zero-padded input."""
        if x.shape[1] == m.in_channels:
            return m(x)
        from ..conv_ops import padded_in_channels
        w = padded_in_channels(m, x.shape[1])
        if isinstance(m, nn.ConvTranspose2d):
            return nn.functional.conv_transpose2d(x, w, m.bias, m.stride, m.padding, m.output_padding, m.groups, m.dilation)
        return nn.functional.conv2d(x, w, m.bias, m.stride, m.padding, m.dilation, m.groups)

    def _cl(self, x):
        return bool(getattr(self, "channels_last", False)) and x.is_cuda and x.dtype == torch.float32

    def _decode(self, lvl, x):
        for i in range(5):
            m = getattr(self, f"conv{lvl}_{i}")
            if self._cl(x) and len(m) == 2:             # conv + LeakyReLU + concatenation as one autograd node (conv_ops)
                from ..conv_ops import dense_conv_cat
                x = dense_conv_cat(m[0], x, float(m[1].negative_slope))
            else:
                x = self._cat((m(x), x))
        return x, self._plain(getattr(self, f"predict_flow{lvl}"), x)

    def forward(self, im1, im2):
        im1 = torch.stack((im1[:, 2], im1[:, 1], im1[:, 0]), 1)        # RGB -> BGR (PWCNet.py:232-233)
        im2 = torch.stack((im2[:, 2], im2[:, 1], im2[:, 0]), 1)
        cl = self._cl(im1)
        if cl:
            # NHWC through every convolution (cuDNN's sm_100 kernels are NHWC-only: with NCHW activations it converts around
            # each of the ~70 convolutions, 1.7 ms of a 5.7 ms closure); the correlation / warp operators take NCHW, so each
            # feature level is converted once
            im1 = im1.contiguous(memory_format=torch.channels_last)
            im2 = im2.contiguous(memory_format=torch.channels_last)
        c1, c2 = self._features(im1), self._features(im2)
        nchw = (lambda t: t.contiguous()) if cl else (lambda t: t)
        corr = self.leakyRELU(self.corr(nchw(c1[6]), nchw(c2[6])))
        x, flow = self._decode(6, self._cat((corr,), pad=True) if cl else corr)
        flows = {6: flow}
        for lvl in (5, 4, 3, 2):
            up_flow = getattr(self, f"deconv{lvl + 1}")(flow)
            up_feat = self._plain(getattr(self, f"upfeat{lvl + 1}"), x)
            warped = self.warp(nchw(c2[lvl]), nchw(up_flow) * _FLOW_SCALE[lvl])
            corr = self.leakyRELU(self.corr(nchw(c1[lvl]), warped))
            x, flow = self._decode(lvl, self._cat((corr, c1[lvl], up_flow, up_feat), pad=True))
            flows[lvl] = flow
        x = self.dc_conv4(self.dc_conv3(self.dc_conv2(self.dc_conv1(x))))
        flow2 = flow + self.dc_conv7(self.dc_conv6(self.dc_conv5(x)))
        if self.training:
            return tuple(20 * self.upsample(f) for f in (flow2, flows[3], flows[4], flows[5], flows[6]))
        return 20 * self.upsample(flow2)
