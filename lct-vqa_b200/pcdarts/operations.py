"""Candidate operations of the PC-DARTS search space — B200 drop-in for
darts_vqa/pcdarts/operations.py (== basic_vqa/pcdarts/operations.py).

Same factory table (`OPS[name](C, stride, affine)`), same module classes, constructor signatures,
sub-module names and parameter registration order (state_dict keys such as `op.1.weight`,
`op.3.running_mean`, `conv_1.weight`, `bn.running_var` are identical), so checkpoints and the
architects' flat `theta` slicing (architect_vqa.py:90-103) carry over.

What differs is execution: on the search path these modules are *parameter containers*.  MixedOp /
Cell / Network (model_search.py) hand their weights to the fused sm_100a kernels of
libpcdarts_sm100.so; no op below is run layer by layer there.  Called on their own, the two
preprocess ops (ReLUConvBN 1x1, FactorizedReduce) go through the same CUDA preprocess kernels the
Cell uses; Identity / Zero are trivial; the remaining stand-alone forwards are not part of the
accelerated path and say so.
"""
import torch
import torch.nn as nn


def _seq(*mods):
    return nn.Sequential(*mods)


def _depthwise(C, k, stride, padding, dilation=1):
    return nn.Conv2d(C, C, kernel_size=k, stride=stride, padding=padding, dilation=dilation, groups=C, bias=False)


def _pointwise(C_in, C_out):
    return nn.Conv2d(C_in, C_out, kernel_size=1, padding=0, bias=False)


class _ContainerOp(nn.Module):
    """An op whose arithmetic lives in the fused kernels of its parent MixedOp."""

    def forward(self, x):
        raise NotImplementedError(
            f"{type(self).__name__} is executed inside the fused MixedOp/Cell CUDA kernels "
            "(libpcdarts_sm100.so); a stand-alone layer-by-layer forward is outside the accelerated path")


class ReLUConvBN(nn.Module):
    """ReLU -> Conv(k, stride, pad) -> BN  (operations.py:22-33).  The search network only uses the
    1x1 / stride 1 / pad 0 form (model_search.py:70-71), which is what the CUDA path implements."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, affine=True):
        super().__init__()
        self.op = _seq(nn.ReLU(inplace=False),
                       nn.Conv2d(C_in, C_out, kernel_size, stride=stride, padding=padding, bias=False),
                       nn.BatchNorm2d(C_out, affine=affine))
        self._spec = (C_in, C_out, kernel_size, stride, padding, affine)

    def forward(self, x):
        from pcd_ops import preprocess_apply
        return preprocess_apply(self, x, fr=False)


class DilConv(_ContainerOp):
    """ReLU -> dilated depthwise -> 1x1 -> BN  (operations.py:35-47)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, affine=True):
        super().__init__()
        self.op = _seq(nn.ReLU(inplace=False),
                       _depthwise(C_in, kernel_size, stride, padding, dilation),
                       _pointwise(C_in, C_out),
                       nn.BatchNorm2d(C_out, affine=affine))


class SepConv(_ContainerOp):
    """(ReLU -> depthwise -> 1x1 -> BN) twice, the first with the stride  (operations.py:50-66)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, affine=True):
        super().__init__()
        self.op = _seq(nn.ReLU(inplace=False),
                       _depthwise(C_in, kernel_size, stride, padding),
                       _pointwise(C_in, C_in),
                       nn.BatchNorm2d(C_in, affine=affine),
                       nn.ReLU(inplace=False),
                       _depthwise(C_in, kernel_size, 1, padding),
                       _pointwise(C_in, C_out),
                       nn.BatchNorm2d(C_out, affine=affine))


class Identity(nn.Module):
    def forward(self, x):
        return x


class Zero(nn.Module):
    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def forward(self, x):
        s = self.stride
        return x.mul(0.) if s == 1 else x[:, :, ::s, ::s].mul(0.)


class FactorizedReduce(nn.Module):
    """ReLU -> two stride-2 1x1 convs (the second on x[:, :, 1:, 1:]) -> concat -> BN (operations.py:90-104)."""

    def __init__(self, C_in, C_out, affine=True):
        super().__init__()
        assert C_out % 2 == 0
        self.relu = nn.ReLU(inplace=False)
        self.conv_1 = nn.Conv2d(C_in, C_out // 2, 1, stride=2, padding=0, bias=False)
        self.conv_2 = nn.Conv2d(C_in, C_out // 2, 1, stride=2, padding=0, bias=False)
        self.bn = nn.BatchNorm2d(C_out, affine=affine)
        self._spec = (C_in, C_out, affine)

    def forward(self, x):
        from pcd_ops import preprocess_apply
        return preprocess_apply(self, x, fr=True)


class _PoolContainer(_ContainerOp):
    """Holds the hyper-parameters of a 3x3 pool candidate; pooled inside the MixedOp kernels."""

    def __init__(self, kind, stride):
        super().__init__()
        self.kind, self.kernel_size, self.stride, self.padding = kind, 3, stride, 1
        self.count_include_pad = False


OPS = {
    'none': lambda C, stride, affine: Zero(stride),
    'avg_pool_3x3': lambda C, stride, affine: _PoolContainer('avg', stride),
    'max_pool_3x3': lambda C, stride, affine: _PoolContainer('max', stride),
    'skip_connect': lambda C, stride, affine: Identity() if stride == 1 else FactorizedReduce(C, C, affine=affine),
    'sep_conv_3x3': lambda C, stride, affine: SepConv(C, C, 3, stride, 1, affine=affine),
    'sep_conv_5x5': lambda C, stride, affine: SepConv(C, C, 5, stride, 2, affine=affine),
    'dil_conv_3x3': lambda C, stride, affine: DilConv(C, C, 3, stride, 2, 2, affine=affine),
    'dil_conv_5x5': lambda C, stride, affine: DilConv(C, C, 5, stride, 4, 2, affine=affine),
}
