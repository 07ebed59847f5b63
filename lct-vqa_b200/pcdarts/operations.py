"""Candidate operations of the PC-DARTS search space — B200 drop-in for
darts_vqa/pcdarts/operations.py (== basic_vqa/pcdarts/operations.py).

Same factory table (`OPS[name](C, stride, affine)`), same module classes, constructor signatures,
sub-module names and parameter registration order (state_dict keys such as `op.1.weight`,
`op.3.running_mean`, `conv_1.weight`, `bn.running_var` are identical), so checkpoints and the
architects' flat `theta` slicing (architect_vqa.py:90-103) carry over.

What differs is execution: on the search path these modules are *parameter containers*.  MixedOp /
Cell / Network (model_search.py) hand their weights to the fused sm_100a kernels of
libpcdarts_sm100.so; no op below is run layer by layer there.  Called on their own — which is how a
network derived from a genotype uses them (pcdarts/model.py) — they run on the library's stand-alone
op kernels (pcd_opmods.py: depthwise / pointwise+BatchNorm / pool kernels on all C channels); the two
preprocess ops (ReLUConvBN 1x1, FactorizedReduce) go through the CUDA preprocess kernels the Cell uses,
followed by the affine kernel when affine=True.  Training-mode BatchNorm only; nothing falls back to
stock torch layers (`stock_forward` exists for the parity tests).
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
    """An op whose arithmetic lives in the fused kernels of its parent MixedOp on the search path and in the stand-alone op
    kernels when it is called on its own."""

    def stock_forward(self, x):
        """The same layers in stock torch — the reference's forward (operations.py:46,65).  Parity tests only."""
        return self.op(x)


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

    def stock_forward(self, x):
        return self.op(x)


class DilConv(_ContainerOp):
    """ReLU -> dilated depthwise -> 1x1 -> BN  (operations.py:35-47)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, affine=True):
        super().__init__()
        self.op = _seq(nn.ReLU(inplace=False),
                       _depthwise(C_in, kernel_size, stride, padding, dilation),
                       _pointwise(C_in, C_out),
                       nn.BatchNorm2d(C_out, affine=affine))

    def forward(self, x):
        from pcd_opmods import unit_apply
        return unit_apply(x, self.op[1], self.op[2], self.op[3], "DilConv")


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

    def forward(self, x):
        from pcd_opmods import unit_apply
        y = unit_apply(x, self.op[1], self.op[2], self.op[3], "SepConv")
        return unit_apply(y, self.op[5], self.op[6], self.op[7], "SepConv")


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

    def stock_forward(self, x):
        x = self.relu(x)
        return self.bn(torch.cat([self.conv_1(x), self.conv_2(x[:, :, 1:, 1:])], dim=1))


class _PoolContainer(_ContainerOp):
    """Holds the hyper-parameters of a 3x3 pool candidate (AvgPool2d(3, stride, 1, count_include_pad=False) /
    MaxPool2d(3, stride, 1), operations.py:6-7); pooled inside the MixedOp kernels on the search path."""

    def __init__(self, kind, stride):
        super().__init__()
        self.kind, self.kernel_size, self.stride, self.padding = kind, 3, stride, 1
        self.count_include_pad = False

    def forward(self, x):
        from pcd_opmods import pool_apply
        return pool_apply(x, self.kind, self.stride)

    def stock_forward(self, x):
        import torch.nn.functional as F
        if self.kind == 'max':
            return F.max_pool2d(x, 3, self.stride, 1)
        return F.avg_pool2d(x, 3, self.stride, 1, count_include_pad=False)


class Conv7x1_1x7(_ContainerOp):
    """ReLU -> Conv(1x7) -> Conv(7x1) -> BN  (operations.py:14-19).  In the reference's OPS table but in no PRIMITIVES list
    (genotypes.py:5-14), so no search or derived network can contain it: kept as a parameter container, no native kernel."""

    def __init__(self, C, stride, affine=True):
        super().__init__()
        self.op = _seq(nn.ReLU(inplace=False),
                       nn.Conv2d(C, C, (1, 7), stride=(1, stride), padding=(0, 3), bias=False),
                       nn.Conv2d(C, C, (7, 1), stride=(stride, 1), padding=(3, 0), bias=False),
                       nn.BatchNorm2d(C, affine=affine))

    def forward(self, x):
        raise RuntimeError("conv_7x1_1x7: no native kernel (PCD_ERR_UNSUPPORTED); it is not in PRIMITIVES")


OPS = {
    'none': lambda C, stride, affine: Zero(stride),
    'avg_pool_3x3': lambda C, stride, affine: _PoolContainer('avg', stride),
    'max_pool_3x3': lambda C, stride, affine: _PoolContainer('max', stride),
    'skip_connect': lambda C, stride, affine: Identity() if stride == 1 else FactorizedReduce(C, C, affine=affine),
    'sep_conv_3x3': lambda C, stride, affine: SepConv(C, C, 3, stride, 1, affine=affine),
    'sep_conv_5x5': lambda C, stride, affine: SepConv(C, C, 5, stride, 2, affine=affine),
    'sep_conv_7x7': lambda C, stride, affine: SepConv(C, C, 7, stride, 3, affine=affine),
    'dil_conv_3x3': lambda C, stride, affine: DilConv(C, C, 3, stride, 2, 2, affine=affine),
    'dil_conv_5x5': lambda C, stride, affine: DilConv(C, C, 5, stride, 4, 2, affine=affine),
    'conv_7x1_1x7': lambda C, stride, affine: Conv7x1_1x7(C, stride, affine),
}
