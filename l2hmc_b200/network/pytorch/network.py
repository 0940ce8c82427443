"""xnet / vnet `LeapfrogLayer`s with the reference's module tree and parameter
names (`network/pytorch/network.py:151-801`, `network/factory.py:21-71`), so a
reference `state_dict` loads unchanged:

    input_layer.{conv_stack.layers.<i>, xlayer, vlayer}, hidden_layers.<i>,
    scale.{coeff, layer}, transf.{coeff, layer}, transl, batch_norm

    z = act(W_x flat(conv?(x)) + W_v flat(v));  z = act(W_i z) ...
    s = a_s e^{c_s} tanh(W_s z),  t = a_t W_t z,  q = a_q e^{c_q} tanh(W_q z)

Where the nets run in bf16 (autocast, BASELINE cfg 5) every dense layer -- input pair, hidden Linears, heads, and
all three GEMMs of their backward passes -- is the hand-written tcgen05 GEMM `l2b_gemm_bf16` (`autograd.TCDense`;
the SU(3) heads and input layer have their own fused kernels on top, `l2b_su3_heads_vupdate`, `l2b_su3_input_layer`).
fp32 / fp64 nets without autocast and the U(1) conv stack go through torch.nn (library calls); everything the
integrator does with (s, t, q) afterwards is in libl2b's fused epilogue kernels.
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ...configs import ConvolutionConfig, InputSpec, NetWeight, NetWeights, NetworkConfig

Tensor = torch.Tensor


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError('l2hmc_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def activation_fn(name: str) -> nn.Module:
    """fresh module per network (reference shares global inplace instances,
    network.py:40-46; the arithmetic is identical)"""
    fns = {
        'elu': lambda: nn.ELU(inplace=True),
        'tanh': lambda: nn.Tanh(),
        'relu': lambda: nn.ReLU(inplace=True),
        'swish': lambda: nn.SiLU(),
        'leaky_relu': lambda: nn.LeakyReLU(inplace=True),
    }
    if name not in fns:
        raise ValueError(f'unknown activation {name!r}')
    return fns[name]()


def flatten(x: Tensor) -> Tensor:
    return x.reshape(x.shape[0], -1)


def nested_children(m: nn.Module) -> dict:
    """module tree as nested dicts, leaves keyed by their class name (network.py:49-57)"""
    children = dict(m.named_children())
    if not children:
        return {m._get_name(): m}
    return {name: nested_children(child) for name, child in children.items()}


def xy_repr(x: Tensor) -> Tensor:
    """angles -> (cos, sin) stacked on a new axis 1 (network.py:65-66)"""
    return torch.stack((torch.cos(x), torch.sin(x)), dim=1)


def init_all(model: nn.Module, init_func: Callable, *params, **kwargs) -> None:
    """apply `init_func(p, *params, **kwargs)` to every parameter (network.py:80-90)"""
    for p in model.parameters():
        init_func(p, *params, **kwargs)


def init_all_by_shape(model: nn.Module, init_funcs: dict) -> None:
    """per-parameter initialiser chosen by the parameter's rank, `init_funcs[str(rank)]`, falling back to
    `init_funcs['default']` (network.py:93-118)"""
    assert 'default' in init_funcs, 'init_funcs must have `default` entry'
    for p in model.parameters():
        init_funcs.get(str(p.dim()), init_funcs['default'])(p)


def init_weights(m: nn.Module, method: str = 'xavier_uniform') -> None:
    """`module.apply`-style initialiser of the `nn.Linear` weights (network.py:121-141)"""
    if not isinstance(m, nn.Linear):
        return
    if method == 'zeros':
        nn.init.zeros_(m.weight)
        nn.init.zeros_(m.bias)
        return
    fn = getattr(nn.init, method if method.endswith('_') else method + '_', None)
    if callable(fn):
        fn(m.weight)


def zero_weights(m: nn.Module) -> None:
    """network.py:144-148"""
    if isinstance(m, nn.Linear):
        nn.init.zeros_(m.weight.data)
        if m.bias is not None:
            nn.init.zeros_(m.bias.data)


def calc_output_size(hw: tuple[int, int], kernel_size, stride: int = 1, pad: int = 0, dilation: int = 1) -> tuple[int, int]:
    """spatial size after a convolution / pooling window (network.py:209-237)"""
    k = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    return tuple((n + 2 * pad - dilation * (kk - 1) - 1) // stride + 1 for n, kk in zip(hw, k))


def dummy_network(inputs: tuple[Tensor, Tensor]) -> tuple[Tensor, Tensor, Tensor]:
    """network.py:69-77"""
    x, _ = inputs
    return torch.zeros_like(x), torch.zeros_like(x), torch.zeros_like(x)



# Parameter `_version`s only move when PYTHON mutates a parameter.  A training step replayed from a CUDA graph
# updates the weights on the device without touching them, so every cache derived from parameter values (the
# bf16 UMMA image of the heads, cast copies, host copies of the step sizes, captured eval graphs) also keys on this
# process-wide generation, which `Trainer.train_step` bumps once per step, eager or replayed.
_WEIGHTS_GENERATION = [0]


def weights_generation() -> int:
    return _WEIGHTS_GENERATION[0]


def bump_weights_generation() -> int:
    _WEIGHTS_GENERATION[0] += 1
    return _WEIGHTS_GENERATION[0]


# A training step captured into a CUDA graph must contain the kernels that derive the bf16 weight images from the
# live weights (the replay moves the weights), but only ONCE per step: `Trainer.train_step` opens a new step token,
# and while a stream is capturing the weight-derived caches are valid exactly within one token.
_STEP_TOKEN = [0]


def new_step_token() -> int:
    _STEP_TOKEN[0] += 1
    return _STEP_TOKEN[0]


def _capture_token() -> int:
    """-1 outside a capture (caches live as long as the weights' versions), else the current step token"""
    return _STEP_TOKEN[0] if torch.cuda.is_current_stream_capturing() else -1


def _tc_tensor_ok(t: Tensor) -> bool:
    """tensors the hand-written kernels can take (the CPU test tier swaps this check together with the `ops` entry
    points, tests/cpu_emulation.py)"""
    return t.is_cuda


def activation_name(a) -> Optional[str]:
    """name of an activation module as the fused kernels know it (None: not one of the reference's, network.py:40-46)"""
    if isinstance(a, nn.Tanh):
        return 'tanh'
    if isinstance(a, nn.ReLU):
        return 'relu'
    if isinstance(a, nn.SiLU):
        return 'swish'
    if isinstance(a, nn.LeakyReLU) and abs(a.negative_slope - 0.01) < 1e-12:
        return 'leaky_relu'
    if isinstance(a, nn.ELU) and abs(a.alpha - 1.0) < 1e-12:
        return 'elu'
    return None


class _WeightImages:
    """bf16 / bf16x3 images of a module's weight matrices for the tensor-core GEMM, cached per weight version
    (and, inside a CUDA-graph capture of a training step, per step: `_capture_token`).  Conv weights are viewed as
    [Cout, Cin n^2]."""

    @staticmethod
    def _as_matrix(w: Tensor, tap_major: bool) -> Tensor:
        """[out, in] view of a Linear / Conv2d weight; `tap_major`: a Conv2d weight [Cout, Cin, n, n] as
        [Cout, (kh, kw, ci)] (channel fastest), the column order of `ops.conv_im2col(..., tap_major=True)`"""
        w = w.detach()
        if tap_major:
            w = w.permute(0, 2, 3, 1)
        return w.reshape(w.shape[0], -1)

    def weight_as_bf16(self, w: Tensor, tap_major: bool = False) -> Tensor:
        """bf16 copy of a weight matrix (what autocast would re-create on every call), cached per weight version;
        re-cast once per step inside a CUDA-graph capture of a training step, where the weights change on every replay"""
        if w.dtype == torch.bfloat16 and not tap_major:
            return w.detach().reshape(w.shape[0], -1)
        cache = self.__dict__.setdefault('_bf16_weights', {})
        key = (w._version, w.data_ptr(), weights_generation(), _capture_token())
        hit = cache.get((id(w), tap_major))
        if hit is None or hit[0] != key:
            hit = (key, self._as_matrix(w, tap_major).to(torch.bfloat16).contiguous())
            cache[(id(w), tap_major)] = hit
        return hit[1]

    def weight_split3(self, w: Tensor, tap_major: bool = False) -> Tensor:
        """bf16x3 split of an fp32 weight matrix ([3, out, in8], ops.split_bf16x3) for the fp32-accurate tensor-core
        GEMM, cached like `weight_as_bf16`"""
        from ... import ops
        cache = self.__dict__.setdefault('_x3_weights', {})
        key = (w._version, w.data_ptr(), weights_generation(), _capture_token())
        hit = cache.get((id(w), tap_major))
        if hit is None or hit[0] != key:
            hit = (key, ops.split_bf16x3(self._as_matrix(w, tap_major).float().contiguous()))
            cache[(id(w), tap_major)] = hit
        return hit[1]


class PeriodicPadding(nn.Module):
    """wraps `size` on BOTH sides of the last two axes (network.py:151-172)"""

    def __init__(self, size: int):
        super().__init__()
        self.size = size

    def forward(self, x: Tensor) -> Tensor:
        assert len(x.shape) >= 3, 'Expected len(x.shape) >= 3'
        x = torch.cat([x[:, :, -self.size:, :], x, x[:, :, 0:self.size, :]], 2)
        return torch.cat([x[:, :, :, -self.size:], x, x[:, :, :, 0:self.size]], 3)


class ScaledTanh(nn.Module):
    """exp(coeff) * tanh(W x + b)   (network.py:175-206)"""

    def __init__(self, in_features: int, out_features: int) -> None:
        super().__init__()
        self.coeff = nn.Parameter(torch.zeros(1, out_features))
        self.layer = nn.Linear(in_features=in_features, out_features=out_features)

    def forward(self, x):
        return self.coeff.exp() * torch.tanh(self.layer(x))


class ConvStack(nn.Module, _WeightImages):
    """network.py:240-346; same `layers` ordering so indices (hence state_dict
    keys) match."""

    def __init__(self, xshape: Sequence[int], conv_config: ConvolutionConfig, activation_fn: Any,
                 use_batch_norm: bool = False) -> None:
        super().__init__()
        if len(xshape) == 3:
            d, nt, nx = xshape[0], xshape[1], xshape[2]
        elif len(xshape) == 4:
            _, d, nt, nx = xshape
        elif len(xshape) == 8:
            d = xshape[1]
            nt, nx = xshape[2], xshape[3]
        else:
            raise ValueError(f'Invalid value for xshape: {xshape}')
        self.d, self.nt, self.nx = d, nt, nx
        self.xshape = xshape
        self.xdim = int(np.prod(xshape[1:]))
        self.activation_fn = activation_fn
        self.layers = nn.ModuleList()
        filters = list(conv_config.filters or [])
        sizes = list(conv_config.sizes or [])
        if len(filters) > 0 and len(filters) == len(sizes):
            self.layers.append(PeriodicPadding(sizes[0] - 1))
            self.layers.append(nn.LazyConv2d(filters[0], sizes[0]))
            for idx, (f, n) in enumerate(zip(filters[1:], sizes[1:])):
                self.layers.append(PeriodicPadding(n - 1))
                self.layers.append(nn.LazyConv2d(f, n))
                if (idx + 1) % 2 == 0:
                    p = 2 if conv_config.pool is None else conv_config.pool[idx]
                    self.layers.append(nn.MaxPool2d(p))
                self.layers.append(self.activation_fn)
        self.layers.append(nn.Flatten())
        if use_batch_norm:
            self.layers.append(nn.LazyBatchNorm1d())
        self.layers.append(nn.LazyLinear(self.xdim))
        self.layers.append(self.activation_fn)

    def tensor_core_mode(self, x: Tensor) -> Optional[str]:
        """'bf16' (autocast / bf16 parameters) or 'x3' (fp32 parameters, the reference's default precision) when the
        stack can run on the hand-written path (gather + tensor-core GEMM + pooling kernels, `_forward_tc`), else
        None: fp64 nets, an activation outside the reference's list, non-square / strided convolutions, lazy layers
        not yet materialised.  `self.tc_conv`: 'auto' (default) | 'never'"""
        if getattr(self, 'tc_conv', 'auto') == 'never' or not _tc_tensor_ok(x) or activation_name(self.activation_fn) is None:
            return None
        convs = [m for m in self.layers if isinstance(m, nn.Conv2d)]
        lins = [m for m in self.layers if isinstance(m, nn.Linear)]
        if not convs or any(isinstance(m, nn.modules.lazy.LazyModuleMixin) for m in convs + lins):
            return None
        for c in convs:
            # (size 1: the reference's PeriodicPadding(0) slices x[:, :, -0:] = all of x and doubles the image)
            if (c.kernel_size[0] < 2 or c.kernel_size[0] != c.kernel_size[1] or tuple(c.stride) != (1, 1) or tuple(c.dilation) != (1, 1)
                    or tuple(c.padding) != (0, 0) or c.groups != 1):
                return None
        if torch.is_autocast_enabled('cuda'):
            return 'bf16' if torch.get_autocast_dtype('cuda') == torch.bfloat16 else None
        dt = convs[0].weight.dtype
        if dt == torch.bfloat16:
            return 'bf16'
        if dt == torch.float32 and x.dtype in (torch.float32, torch.bfloat16):
            return 'x3'
        return None

    def _forward_tc(self, x: Tensor, mode: str) -> Tensor:
        """the same layer list, block by block, on the hand-written kernels; activations travel NHWC"""
        from ... import autograd as ag
        act = activation_name(self.activation_fn)
        layers = list(self.layers)
        nchw, i = True, 0
        # `self.conv_precision` of fp32 stacks: 'tf32' (default: bf16x2 operands, 16 mantissa bits -- the reference's
        # own GPU path runs these convolutions on cuDNN with TF32, 10 bits, `torch.backends.cudnn.allow_tf32`; two
        # thirds of the gathered bytes and half the products of the exact mode) | 'fp32' (bf16x3 operands,
        # fp32-accurate: parity with the reference's CPU path to 1e-4 in every gradient).  The Linear at the end of
        # the stack is fp32-accurate either way, as torch's fp32 matmul is.
        cmode = 'x2' if (mode == 'x3' and getattr(self, 'conv_precision', 'tf32') == 'tf32') else mode
        while i < len(layers):
            m = layers[i]
            if isinstance(m, PeriodicPadding):
                conv = layers[i + 1]
                assert isinstance(conv, nn.Conv2d) and m.size == conv.kernel_size[0] - 1
                nxt = layers[i + 2] if i + 2 < len(layers) else None
                if isinstance(nxt, nn.MaxPool2d):
                    x = ag.ConvPeriodic.apply(x, conv.weight, conv.bias, nchw, cmode, None, self)
                    p = nxt.kernel_size if isinstance(nxt.kernel_size, int) else nxt.kernel_size[0]
                    x = ag.PoolAct.apply(x, int(p), act)
                    i += 4                                   # pad, conv, pool, activation
                elif nxt is self.activation_fn:
                    x = ag.ConvPeriodic.apply(x, conv.weight, conv.bias, nchw, cmode, act, self)
                    i += 3
                else:                                        # the first block: no activation (network.py:296-307)
                    x = ag.ConvPeriodic.apply(x, conv.weight, conv.bias, nchw, cmode, None, self)
                    i += 2
                nchw = False
            elif isinstance(m, nn.Flatten):
                if not nchw:                                 # the Linear's columns are in NCHW order
                    x = x.permute(0, 3, 1, 2)
                x = x.reshape(x.shape[0], -1)
                i += 1
            elif isinstance(m, nn.Linear):
                fused = act if (i + 1 < len(layers) and layers[i + 1] is self.activation_fn) else None
                x = ag.TCDense.apply(fused, self, mode, x, m.weight, m.bias)
                i += 2 if fused is not None else 1
            else:                                            # batch norm
                x = m(x.float() if mode == 'x3' else x)
                i += 1
        return x

    def forward(self, x: Tensor) -> Tensor:
        if tuple(x.shape) != tuple(self.xshape):
            try:
                x = x.reshape(x.shape[0], self.d + 2, self.nt, self.nx)
            except (ValueError, RuntimeError):
                x = x.reshape((x.shape[0], *self.xshape[1:]))
        mode = self.tensor_core_mode(x)
        if mode is not None:
            return self._forward_tc(x, mode)
        for layer in self.layers:
            x = layer(x)
        return x


class InputLayer(nn.Module):
    """network.py:349-451"""

    def __init__(self, xshape: Sequence[int], network_config: NetworkConfig, activation_fn: Callable,
                 conv_config: Optional[ConvolutionConfig] = None) -> None:
        super().__init__()
        self.xshape = xshape
        self.activation_fn = activation_fn
        conv_stack: nn.Module = nn.Identity()
        if conv_config is not None and conv_config.filters is not None and len(conv_config.filters) > 0:
            conv_stack = ConvStack(xshape=xshape, conv_config=conv_config, activation_fn=activation_fn)
        self.conv_stack = conv_stack
        self.xlayer = nn.LazyLinear(network_config.units[0])
        self.vlayer = nn.LazyLinear(network_config.units[0])

    def forward(self, inputs: tuple[Tensor, Tensor]) -> Tensor:
        x, v = inputs
        x = self.conv_stack(x)
        v = self.vlayer(flatten(v))
        x = self.xlayer(flatten(x))
        return self.activation_fn(x + v)


class LeapfrogLayer(nn.Module, _WeightImages):
    """network.py:454-551"""

    def __init__(self, xshape: Sequence[int], network_config: NetworkConfig,
                 input_shapes: Optional[dict] = None, net_weight: Optional[NetWeight] = None,
                 conv_config: Optional[ConvolutionConfig] = None, name: Optional[str] = None):
        super().__init__()
        self.xshape = xshape
        self.nw = NetWeight(1., 1., 1.) if net_weight is None else net_weight
        self.net_config = network_config
        self.name = name if name is not None else 'network'
        self.xdim = int(np.prod(xshape[1:]))
        act = network_config.activation_fn
        self.activation_fn = activation_fn(act) if isinstance(act, str) else act
        self.input_layer = InputLayer(xshape=xshape, network_config=network_config,
                                      activation_fn=self.activation_fn, conv_config=conv_config)
        self.units = list(network_config.units)
        self.hidden_layers = nn.ModuleList(
            [nn.Linear(self.units[i], u) for i, u in enumerate(self.units[1:])])
        self.scale = ScaledTanh(self.units[-1], self.xdim)
        self.transf = ScaledTanh(self.units[-1], self.xdim)
        self.transl = nn.Linear(self.units[-1], self.xdim)
        self.dropout = nn.Dropout(network_config.dropout_prob)
        if network_config.use_batch_norm:
            self.batch_norm = nn.BatchNorm1d(self.units[-1])

    def set_net_weight(self, net_weight: NetWeight):
        self.nw = net_weight

    # ---- dense layers on the tensor-core GEMM (csrc/l2b_gemm.cu) -----------------------------------------------
    def tensor_core_dense(self, *inputs: Tensor) -> Optional[str]:
        """how the dense layers run on the hand-written tensor-core GEMM (l2b_gemm_bf16), or None (library path):
        'bf16' whenever the nets run in bf16 anyway (autocast, BASELINE cfg 5, or bf16 parameters); 'x3' for fp32
        parameters without autocast (the reference's default precision): bf16x3 operand splits, fp32-accurate.
        fp64 nets and activations outside the reference's list (network.py:40-46) stay on torch.
        `self.tc_dense`: 'auto' (default) | 'never'"""
        if getattr(self, 'tc_dense', 'auto') == 'never' or self.input_activation_name() is None:
            return None
        if not all(_tc_tensor_ok(t) for t in inputs):
            return None
        if isinstance(self.input_layer.xlayer, nn.modules.lazy.LazyModuleMixin):
            return None                     # the materialising dummy call (network.py:572-631) goes through torch
        if torch.is_autocast_enabled('cuda'):
            return 'bf16' if torch.get_autocast_dtype('cuda') == torch.bfloat16 else None
        dt = self.transl.weight.dtype
        if dt == torch.bfloat16:
            return 'bf16'
        if dt == torch.float32 and all(t.dtype in (torch.float32, torch.bfloat16) for t in inputs):
            return 'x3'
        return None

    def _hidden_tc(self, inputs: tuple[Tensor, Tensor], mode: str) -> Tensor:
        from ... import autograd as ag
        il, act = self.input_layer, self.input_activation_name()
        x, v = inputs
        z = ag.TCDense.apply(act, self, mode, flatten(x), il.xlayer.weight, il.xlayer.bias,
                             flatten(v), il.vlayer.weight, il.vlayer.bias)
        for layer in self.hidden_layers:
            z = ag.TCDense.apply(act, self, mode, z, layer.weight, layer.bias)
        return z

    def hidden(self, inputs: tuple[Tensor, Tensor]) -> Tensor:
        """everything in front of the three output heads (network.py:536-545)"""
        mode = self.tensor_core_dense(*inputs)
        if mode is not None and not self.dense_input():
            cs = self.input_layer.conv_stack             # U(1) xnet: the conv stack on its own hand-written path
            mode = mode if cs.tensor_core_mode(inputs[0]) == mode else None
            if mode is not None:
                inputs = (cs(inputs[0]), inputs[1])
        if mode is not None:
            z = self._hidden_tc(inputs, mode)
        else:
            z = self.input_layer(inputs)
            for layer in self.hidden_layers:
                z = self.activation_fn(layer(z))
        if self.net_config.dropout_prob > 0:
            z = self.dropout(z)
        if self.net_config.use_batch_norm:
            z = self.batch_norm(z)
        return z

    def hidden_from_pre(self, pre: Tensor) -> Tensor:
        """`hidden` given the input layer's pre-activation sum W_x f(x) + b_x + W_v v + b_v (computed by
        `ops.u1_input_layer` in one pass over the fields)"""
        z = self.activation_fn(pre)
        for layer in self.hidden_layers:
            z = self.activation_fn(layer(z))
        if self.net_config.dropout_prob > 0:
            z = self.dropout(z)
        if self.net_config.use_batch_norm:
            z = self.batch_norm(z)
        return z

    def hidden_tail(self, z: Tensor) -> Tensor:
        """`hidden` from the output of the input layer on (network.py:538-545): the remaining hidden Linears,
        dropout, batch norm"""
        mode = self.tensor_core_dense(z) if self.hidden_layers else None
        if mode is not None:
            from ... import autograd as ag
            for layer in self.hidden_layers:
                z = ag.TCDense.apply(self.input_activation_name(), self, mode, z, layer.weight, layer.bias)
        else:
            for layer in self.hidden_layers:
                z = self.activation_fn(layer(z))
        if self.net_config.dropout_prob > 0:
            z = self.dropout(z)
        if self.net_config.use_batch_norm:
            z = self.batch_norm(z)
        return z

    def input_activation_name(self) -> Optional[str]:
        """name of the input layer's activation as the fused kernel knows it (None: not one of them)"""
        return activation_name(self.activation_fn)

    def input_pack(self):
        """bf16 K-major image of the two input Linears for the tensor-core input layer (ops.su3_input_layer);
        rebuilt only when one of the four parameters changed (same keys as `heads_pack`)"""
        from ... import ops
        il = self.input_layer
        ps = (il.xlayer.weight, il.xlayer.bias, il.vlayer.weight, il.vlayer.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps) + (weights_generation(),)
        cached = getattr(self, '_input_pack', None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                pack = ops.su3_input_pack(ps[0], ps[2], ps[1], ps[3], self.input_activation_name())
            cached = (key, pack)
            self._input_pack = cached
        return cached[1]

    def dense_input(self) -> bool:
        """no conv stack in front of the input Linears"""
        return isinstance(self.input_layer.conv_stack, nn.Identity)

    def heads(self, z: Tensor) -> tuple[Tensor, Tensor, Tensor]:
        """network.py:546-548"""
        mode = self.tensor_core_dense(z)
        if mode is not None:
            from ... import autograd as ag
            s = self.nw.s * (self.scale.coeff.exp() * ag.TCDense.apply('tanh', self, mode, z, self.scale.layer.weight,
                                                                       self.scale.layer.bias))
            t = self.nw.t * ag.TCDense.apply(None, self, mode, z, self.transl.weight, self.transl.bias)
            q = self.nw.q * (self.transf.coeff.exp() * ag.TCDense.apply('tanh', self, mode, z, self.transf.layer.weight,
                                                                        self.transf.layer.bias))
            return s, t, q
        s = self.nw.s * self.scale(z)
        t = self.nw.t * self.transl(z)
        q = self.nw.q * self.transf(z)
        return s, t, q

    def forward(self, inputs: tuple[Tensor, Tensor]) -> tuple[Tensor, Tensor, Tensor]:
        return self.heads(self.hidden(inputs))

    def head_params(self) -> tuple[Tensor, ...]:
        return (self.scale.layer.weight, self.scale.layer.bias, self.scale.coeff,
                self.transl.weight, self.transl.bias,
                self.transf.layer.weight, self.transf.layer.bias, self.transf.coeff)

    def head_weights_as(self, dtype: torch.dtype) -> tuple[Tensor, Tensor, Tensor]:
        """(W_s, W_t, W_q) in `dtype` for the Linear backward (dz = g W), cast once per weight version"""
        ws, _, _, wt, _, wq, _, _ = self.head_params()
        if ws.dtype == dtype:
            return ws.detach(), wt.detach(), wq.detach()
        key = (dtype, ws._version, wt._version, wq._version, ws.data_ptr(), weights_generation(), _capture_token())
        cached = getattr(self, '_head_weights_cast', None)
        if cached is None or cached[0] != key:
            cached = (key, tuple(w.detach().to(dtype) for w in (ws, wt, wq)))
            self._head_weights_cast = cached
        return cached[1]

    def heads_pack(self, perm: Optional[Tensor] = None):
        """bf16 UMMA tile image of the three head matrices + epilogue constants for the fused
        tcgen05 kernel (ops.su3_heads_vupdate); rebuilt only when a head parameter or the net
        weights changed (optimizer step, load_state_dict, set_net_weight).  `perm` (a LongTensor
        over the xdim outputs) packs the rows in another output order -- the planar field layout --
        so that the kernel writes that layout directly; cached separately."""
        from ... import ops
        ps = self.head_params()
        key = tuple((p.data_ptr(), p._version) for p in ps) + (self.nw.s, self.nw.t, self.nw.q, weights_generation())
        if self.training and any(p.requires_grad for p in ps):
            # a training step captured in a CUDA graph changes the weights on every replay without Python running:
            # the pack kernel must then be part of the graph -- once per step (`_capture_token`)
            key = key + (_capture_token(),)
        if perm is not None:
            cached = getattr(self, '_heads_pack_perm', None)
            if cached is None or cached[0] != key or cached[2] is not perm:
                ws, bs, cs, wt, bt, wq, bq, cq = ps
                with torch.no_grad():
                    sel = lambda a, d: a.detach().index_select(d, perm)  # noqa: E731
                    pack = ops.vnet_pack_heads(sel(ws, 0), sel(wt, 0), sel(wq, 0), sel(bs, 0), sel(bt, 0), sel(bq, 0),
                                               sel(cs, 1), sel(cq, 1), self.nw.s, self.nw.t, self.nw.q)
                cached = (key, pack, perm)
                self._heads_pack_perm = cached
            return cached[1]
        cached = getattr(self, '_heads_pack', None)
        if cached is None or cached[0] != key:
            ws, bs, cs, wt, bt, wq, bq, cq = ps
            with torch.no_grad():
                pack = ops.vnet_pack_heads(ws, wt, wq, bs, bt, bq, cs, cq, self.nw.s, self.nw.t, self.nw.q)
            cached = (key, pack)
            self._heads_pack = cached
        return cached[1]


def get_network(xshape: Sequence[int], network_config: NetworkConfig, input_shapes: Optional[dict] = None,
                net_weight: Optional[NetWeight] = None, conv_config: Optional[ConvolutionConfig] = None,
                name: Optional[str] = None) -> LeapfrogLayer:
    """a LeapfrogLayer with its lazy layers still unmaterialised (network.py:554-569)"""
    return LeapfrogLayer(xshape=xshape, network_config=network_config, input_shapes=input_shapes,
                         net_weight=net_weight, conv_config=conv_config, name=name)


def get_and_call_network(xshape: Sequence[int], *, network_config: NetworkConfig, is_xnet: bool, group,
                         input_shapes=None, net_weight=None, conv_config=None, name=None) -> LeapfrogLayer:
    """Build a LeapfrogLayer on the GPU and materialise its lazy layers with one
    dummy call, as the reference does (network.py:572-631); only the SHAPES of
    the dummy inputs matter."""
    dev = _device()
    net = LeapfrogLayer(xshape=xshape, network_config=network_config, input_shapes=input_shapes,
                        net_weight=net_weight, conv_config=conv_config, name=name).to(dev)
    nb = 2
    dt = torch.get_default_dtype()
    gname = getattr(group, '_name', None)
    if gname == 'SU3':
        lat = tuple(xshape[1:6])                       # (4, T, X, Y, Z)
        if is_xnet:                                    # cat(real, imag) on dim 1 (network.py:614-616)
            x = torch.zeros((nb, 2 * lat[0], *lat[1:], 3, 3), dtype=dt, device=dev)
            v = torch.zeros_like(x)
        else:                                          # group_to_vec -> 8 reals per link
            x = torch.zeros((nb, *lat, 8), dtype=dt, device=dev)
            v = torch.zeros_like(x)
    else:
        d, T, X = xshape[1], xshape[2], xshape[3]
        x = torch.zeros((nb, 2 * d if is_xnet else d, T, X), dtype=dt, device=dev)
        v = torch.zeros((nb, d * T * X), dtype=dt, device=dev)
    was_training = net.training
    net.eval()                                          # keep BatchNorm statistics untouched
    with torch.no_grad():
        _ = net((x, v))
    net.train(was_training)
    return net


class NetworkFactory:
    """network/factory.py:21-71 + network.py:634-801"""

    def __init__(self, input_spec: InputSpec, network_config: NetworkConfig,
                 conv_config: Optional[ConvolutionConfig] = None, net_weights: Optional[NetWeights] = None,
                 build_unused_su3_xnet: bool = True):
        if net_weights is None:
            net_weights = NetWeights(x=NetWeight(1., 1., 1.), v=NetWeight(1., 1., 1.))
        self.nw = net_weights
        self.input_spec = input_spec
        self.network_config = network_config
        self.conv_config = conv_config
        # The SU(3) x-update never calls xnet (dynamics.py:1420-1425) but the
        # reference still builds it; keep it by default so state_dicts match.
        self.build_unused_su3_xnet = build_unused_su3_xnet
        self.config = {'net_weights': self.nw, 'input_spec': self.input_spec, 'network_config': self.network_config}

    def get_build_configs(self):
        return {
            'xnet': {'net_weight': self.nw.x, 'xshape': self.input_spec.xshape,
                     'input_shapes': self.input_spec.xnet, 'network_config': self.network_config,
                     'conv_config': self.conv_config},
            'vnet': {'net_weight': self.nw.v, 'xshape': self.input_spec.xshape,
                     'input_shapes': self.input_spec.vnet, 'network_config': self.network_config},
        }

    def build_xnet(self, group, name: Optional[str] = None) -> nn.Module:
        if getattr(group, '_name', None) == 'SU3' and not self.build_unused_su3_xnet:
            return nn.Identity()
        return get_and_call_network(
            xshape=self.input_spec.xshape, network_config=self.network_config, is_xnet=True, group=group,
            input_shapes=self.input_spec.xnet, net_weight=self.nw.x, conv_config=self.conv_config,
            name='xnet' if name is None else f'xnet/{name}')

    def build_vnet(self, group, name: Optional[str] = None) -> LeapfrogLayer:
        return get_and_call_network(
            xshape=self.input_spec.xshape, network_config=self.network_config, is_xnet=False, group=group,
            input_shapes=self.input_spec.vnet, net_weight=self.nw.v, conv_config=self.conv_config,
            name='vnet' if name is None else f'vnet/{name}')

    def build_networks(self, n: int, split_xnets: bool, group) -> nn.ModuleDict:
        assert n >= 1, 'Must build at least one network'
        if n == 1:
            return nn.ModuleDict({'xnet': self.build_xnet(group=group), 'vnet': self.build_vnet(group=group)})
        vnet, xnet = nn.ModuleDict(), nn.ModuleDict()
        for lf in range(n):
            vnet[f'{lf}'] = self.build_vnet(group=group, name=f'{lf}')
            if split_xnets:
                xnet[f'{lf}'] = nn.ModuleDict({
                    'first': self.build_xnet(group=group, name=f'{lf}/first'),
                    'second': self.build_xnet(group=group, name=f'{lf}/second'),
                })
            else:
                xnet[f'{lf}'] = self.build_xnet(group=group, name=f'{lf}')
        return nn.ModuleDict({'xnet': xnet, 'vnet': vnet})
