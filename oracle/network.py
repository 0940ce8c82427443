"""oracle/network.py -- TEST INFRASTRUCTURE (the checker), never the product path.

Plain-numpy restatement of the reference's `LeapfrogLayer` forward pass in
`.eval()` mode (`network/pytorch/network.py:454-551`), driven by a flat
``{name: ndarray}`` dict with the reference's own `state_dict` key names:

    input_layer.conv_stack.layers.<i>.{weight,bias}   (Conv2d / final Linear)
    input_layer.{xlayer,vlayer}.{weight,bias}
    hidden_layers.<i>.{weight,bias}
    batch_norm.{weight,bias,running_mean,running_var}
    scale.{coeff,layer.weight,layer.bias}  transf.{...}  transl.{weight,bias}
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np


def _act(name: str):
    """ACTIVATION_FNS (network.py:40-46)"""
    if name == 'tanh':
        return np.tanh
    if name == 'relu':
        return lambda z: np.maximum(z, 0)
    if name == 'leaky_relu':
        return lambda z: np.where(z >= 0, z, z * np.asarray(0.01, z.dtype))
    if name == 'elu':
        return lambda z: np.where(z > 0, z, np.expm1(np.minimum(z, 0)))
    if name == 'swish':
        return lambda z: z / (1 + np.exp(-z))
    raise ValueError(name)


def linear(z, w, b=None):
    y = z @ w.T
    return y if b is None else y + b


def periodic_pad(x, size: int):
    """PeriodicPadding (network.py:151-172): wraps `size` on BOTH sides of the
    last two axes."""
    x = np.concatenate([x[:, :, -size:, :], x, x[:, :, :size, :]], axis=2)
    return np.concatenate([x[:, :, :, -size:], x, x[:, :, :, :size]], axis=3)


def conv2d(x, w, b):
    """valid cross-correlation, stride 1 (nn.Conv2d defaults)"""
    k0, k1 = w.shape[2], w.shape[3]
    win = np.lib.stride_tricks.sliding_window_view(x, (k0, k1), axis=(2, 3))
    # win: [nb, cin, H', W', k0, k1]
    y = np.einsum('bchwij,ocij->bohw', win, w, optimize=True)
    return y + b[None, :, None, None]


def maxpool2d(x, p: int):
    nb, c, h, w = x.shape
    h2, w2 = h // p, w // p
    x = x[:, :, :h2 * p, :w2 * p].reshape(nb, c, h2, p, w2, p)
    return x.max(axis=(3, 5))


def conv_stack(x, sd: dict, prefix: str, filters: Sequence[int],
               sizes: Sequence[int], pool: Optional[Sequence[int]], act):
    """ConvStack.forward (network.py:240-346).  Layer indices inside
    `layers` follow the reference's append order."""
    li = 0

    def nxt():
        nonlocal li
        i = li
        li += 1
        return i
    x = periodic_pad(x, sizes[0] - 1)
    nxt()
    i = nxt()
    x = conv2d(x, sd[f'{prefix}.layers.{i}.weight'], sd[f'{prefix}.layers.{i}.bias'])
    for idx, (_f, n) in enumerate(zip(filters[1:], sizes[1:])):
        x = periodic_pad(x, n - 1)
        nxt()
        i = nxt()
        x = conv2d(x, sd[f'{prefix}.layers.{i}.weight'], sd[f'{prefix}.layers.{i}.bias'])
        if (idx + 1) % 2 == 0:
            p = 2 if pool is None else pool[idx]
            x = maxpool2d(x, p)
            nxt()
        x = act(x)
        nxt()
    x = x.reshape(x.shape[0], -1)
    nxt()  # Flatten
    i = nxt()
    x = linear(x, sd[f'{prefix}.layers.{i}.weight'], sd[f'{prefix}.layers.{i}.bias'])
    return act(x)


def leapfrog_layer(x, v, sd: dict, *, activation: str, net_weight=(1., 1., 1.),
                   use_batch_norm: bool = False, conv: Optional[dict] = None,
                   conv_in_shape: Optional[Sequence[int]] = None):
    """LeapfrogLayer.forward in eval mode -> (s, t, q), each [nb, xdim].

    `conv`: dict(filters=, sizes=, pool=) or None; `conv_in_shape` is the
    [C, T, X] the conv stack reshapes `x` to (network.py:334-340)."""
    act = _act(activation)
    nb = x.shape[0]
    if conv is not None and conv.get('filters'):
        xin = x.reshape(nb, *conv_in_shape)
        xz = conv_stack(xin, sd, 'input_layer.conv_stack', conv['filters'],
                        conv['sizes'], conv.get('pool'), act)
    else:
        xz = x.reshape(nb, -1)
    vz = linear(v.reshape(nb, -1), sd['input_layer.vlayer.weight'],
                sd['input_layer.vlayer.bias'])
    xz = linear(xz, sd['input_layer.xlayer.weight'], sd['input_layer.xlayer.bias'])
    z = act(xz + vz)
    i = 0
    while f'hidden_layers.{i}.weight' in sd:
        z = act(linear(z, sd[f'hidden_layers.{i}.weight'], sd[f'hidden_layers.{i}.bias']))
        i += 1
    # dropout is identity in eval mode (network.py:542-543)
    if use_batch_norm:
        z = ((z - sd['batch_norm.running_mean'])
             / np.sqrt(sd['batch_norm.running_var'] + np.asarray(1e-5, z.dtype))
             * sd['batch_norm.weight'] + sd['batch_norm.bias'])
    ws, wt, wq = (np.asarray(w, dtype=z.dtype) for w in net_weight)
    s = ws * np.exp(sd['scale.coeff']) * np.tanh(
        linear(z, sd['scale.layer.weight'], sd['scale.layer.bias']))
    t = wt * linear(z, sd['transl.weight'], sd['transl.bias'])
    q = wq * np.exp(sd['transf.coeff']) * np.tanh(
        linear(z, sd['transf.layer.weight'], sd['transf.layer.bias']))
    return s, t, q


def sub_state_dict(sd: dict, prefix: str) -> dict:
    """select `prefix.` keys of a Dynamics state_dict and strip the prefix"""
    p = prefix + '.'
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}
