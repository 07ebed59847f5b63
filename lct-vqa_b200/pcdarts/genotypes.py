"""Search-space vocabulary (reference: darts_vqa/pcdarts/genotypes.py:3-14).

The order of PRIMITIVES defines the columns of alphas_normal / alphas_reduce and the index of every
candidate op inside MixedOp._ops, so it is part of the parameter-naming contract.
"""
from collections import namedtuple

Genotype = namedtuple('Genotype', 'normal normal_concat reduce reduce_concat')

PRIMITIVES = [
    'none',
    'max_pool_3x3',
    'avg_pool_3x3',
    'skip_connect',
    'sep_conv_3x3',
    'sep_conv_5x5',
    'dil_conv_3x3',
    'dil_conv_5x5',
]
