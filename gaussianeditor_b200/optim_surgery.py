"""Optimizer-state surgery for a Gaussian cloud whose size changes while it trains (densify / prune / delete).

Behaviour of /root/reference/gaussiansplatting/scene/gaussian_model.py:553-641: every param group holds exactly one
tensor whose first dimension is the Gaussian index, and is identified by ``group["name"]``. When rows are removed or
appended the parameter object must be replaced (its shape changes), and Adam's moment estimates must follow the rows:
kept rows keep their ``exp_avg`` / ``exp_avg_sq``, new rows start from zero, ``step`` is untouched. Works with any
``torch.optim`` optimizer whose per-parameter state tensors are row-aligned with the parameter (Adam, AdamW, RMSprop,
SGD momentum); scalar / 0-dim state entries (``step``) are carried over unchanged.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
import torch.nn as nn


def _swap(optimizer, group, new_value: torch.Tensor, row_op: Callable[[torch.Tensor], torch.Tensor]) -> nn.Parameter:
    old = group["params"][0]
    state = optimizer.state.pop(old, None)
    new = nn.Parameter(new_value.requires_grad_(True))
    group["params"][0] = new
    if state is not None:
        for k, v in list(state.items()):
            if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == old.shape[0]:
                state[k] = row_op(v)
        optimizer.state[new] = state
    return new


def _single(group):
    if len(group["params"]) != 1:
        raise ValueError("each param group must hold exactly one tensor (gaussian_model.py:604)")
    return group["params"][0]


def prune_optimizer(optimizer, keep_mask: torch.Tensor) -> Dict[str, nn.Parameter]:
    """Keep the rows where `keep_mask` is True in every group (``_prune_optimizer``, :568-589)."""
    out = {}
    for group in optimizer.param_groups:
        p = _single(group)
        out[group["name"]] = _swap(optimizer, group, p.detach()[keep_mask], lambda v: v[keep_mask])
    return out


def cat_tensors_to_optimizer(optimizer, extension: Dict[str, torch.Tensor]) -> Dict[str, nn.Parameter]:
    """Append ``extension[name]`` rows to every group; their moments start at zero (:603-641)."""
    out = {}
    for group in optimizer.param_groups:
        p = _single(group)
        ext = extension[group["name"]]
        if ext.shape[1:] != p.shape[1:]:
            raise ValueError(f"extension of '{group['name']}' has shape {tuple(ext.shape)}, expected [*,{tuple(p.shape[1:])}]")
        ext = ext.detach().to(device=p.device, dtype=p.dtype)
        out[group["name"]] = _swap(optimizer, group, torch.cat((p.detach(), ext), dim=0),
                                   lambda v: torch.cat((v, torch.zeros((ext.shape[0],) + tuple(v.shape[1:]),
                                                                       dtype=v.dtype, device=v.device)), dim=0))
    return out


def replace_tensor_to_optimizer(optimizer, tensor: torch.Tensor, name: str) -> Dict[str, nn.Parameter]:
    """Replace the values of group `name` and reset its moments to zero (``reset_opacity`` path, :553-566)."""
    out = {}
    for group in optimizer.param_groups:
        if group["name"] != name:
            continue
        p = _single(group)
        if tensor.shape != p.shape:
            raise ValueError("replacement must keep the shape")
        out[name] = _swap(optimizer, group, tensor.detach().clone(), torch.zeros_like)
    return out
