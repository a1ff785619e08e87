"""GMA attention pieces on B200 (row f-2): the row softmax of the 7040 x 7040 similarity in fp16 storage with fp32
arithmetic (csrc/attention.cu), as one pass forward and one pass backward.  Reference: models/gma/gma.py:54-76."""
from __future__ import annotations

import torch

from . import _lib


class _SoftmaxRowsF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim):
        lib = _lib.load()
        shape = sim.shape
        sim2 = sim.contiguous().view(-1, shape[-1])
        attn = torch.empty_like(sim2)
        _lib.check(lib.pcfa_softmax_rows_f16_forward(_lib.ptr(sim2), _lib.ptr(attn), sim2.shape[0], sim2.shape[1], _lib.stream()),
                   "pcfa_softmax_rows_f16_forward")
        ctx.save_for_backward(attn)
        return attn.view(shape)

    @staticmethod
    def backward(ctx, gattn):
        lib = _lib.load()
        (attn,) = ctx.saved_tensors
        g = gattn.to(torch.float16).contiguous().view(attn.shape)
        gsim = torch.empty_like(attn)
        _lib.check(lib.pcfa_softmax_rows_f16_backward(_lib.ptr(attn), _lib.ptr(g), _lib.ptr(gsim), attn.shape[0], attn.shape[1],
                                                      _lib.stream()), "pcfa_softmax_rows_f16_backward")
        return gsim.view(gattn.shape)


def softmax_rows_supported(sim: torch.Tensor) -> bool:
    return (sim.is_cuda and sim.dtype == torch.float16 and sim.shape[-1] % 8 == 0 and sim.shape[-1] <= 16384
            and sim.numel() > 0)


def softmax_rows_f16(sim: torch.Tensor) -> torch.Tensor:
    """softmax over the last dimension of an fp16 CUDA tensor; fp16 result (fp32 arithmetic inside)."""
    if not softmax_rows_supported(sim):
        raise RuntimeError("softmax_rows_f16: fp16 CUDA tensor with last dim % 8 == 0 and <= 16384 expected "
                           "(pcfa_b200 has no CPU path)")
    return _SoftmaxRowsF16.apply(sim)
