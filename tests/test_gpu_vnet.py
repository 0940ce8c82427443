"""GPU tier (-m gpu): the tcgen05 heads GEMM fused with the momentum update
(l2b_su3_heads_vupdate) against a float64 torch evaluation of the same bf16-rounded
operands (network.py:536-548 + dynamics.py:1266-1297)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _case(nb, xdim, hidden, seed, wdtype=torch.float32):
    g = torch.Generator(device='cpu').manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)  # noqa: E731
    w = [(r(xdim, hidden) / hidden ** 0.5).to(wdtype).to(DEV) for _ in range(3)]
    b = [(0.1 * r(xdim)).to(wdtype).to(DEV) for _ in range(3)]
    cs, cq = (0.1 * r(1, xdim)).to(wdtype).to(DEV), (0.1 * r(1, xdim)).to(wdtype).to(DEV)
    z = torch.tanh(r(nb, hidden)).to(torch.bfloat16).to(DEV)
    v = torch.complex(r(nb, xdim), r(nb, xdim)).to(DEV)
    f = torch.complex(r(nb, xdim), r(nb, xdim)).to(DEV)
    return w, b, cs, cq, z, v, f


def _reference(w, b, cs, cq, nw, z, v, f, eps, sign):
    """float64 evaluation with the operands rounded to bf16 exactly as the kernel sees them"""
    zz = z.double()
    wr = [x.to(torch.bfloat16).double() for x in w]
    s = nw[0] * cs.double().exp() * torch.tanh(zz @ wr[0].T + b[0].double())
    t = nw[1] * (zz @ wr[1].T + b[1].double())
    q = nw[2] * cq.double().exp() * torch.tanh(zz @ wr[2].T + b[2].double())
    lj = sign * eps * s / 2
    if sign > 0:
        out = lj.exp() * v - 0.5 * eps * (f * (eps * q).exp() + t)
    else:
        out = lj.exp() * (v + 0.5 * eps * (f * (eps * q).exp() + t))
    return out, lj.sum(1), torch.stack([s, t, q])


@pytest.mark.parametrize('nb,xdim,hidden', [(64, 1152, 256), (5, 256, 64), (70, 4320, 128), (130, 1000 * 9, 40)])
@pytest.mark.parametrize('sign', [+1, -1])
def test_heads_vupdate_matches_float64_reference(nb, xdim, hidden, sign):
    from l2hmc_b200 import ops
    w, b, cs, cq, z, v, f = _case(nb, xdim, hidden, seed=nb + hidden)
    nw = (0.7, 1.3, 0.9)
    pack = ops.vnet_pack_heads(w[0], w[1], w[2], b[0], b[1], b[2], cs, cq, *nw)
    eps = 0.11
    out, logdet, stq = ops.su3_heads_vupdate(z, pack, v, f, eps, sign, want_stq=True)
    want, wld, wstq = _reference(w, b, cs, cq, nw, z, v, f, eps, sign)
    # fp32 accumulation / tanhf / expf on bf16 operands: 1e-5 relative (BASELINE north-star fp32 tolerance)
    assert float((stq.double() - wstq).abs().max()) < 2e-5 * max(1.0, float(wstq.abs().max()))
    assert float((out - want).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
    assert float((logdet - wld).abs().max()) < 1e-5 * max(1.0, float(wld.abs().max()))
    # interior tiles take the software-pipelined epilogue when no s/t/q dump is requested: same arithmetic per
    # element (bit-equal v'), another fixed grouping of the fp32 logdet partial sums
    out2, logdet2 = ops.su3_heads_vupdate(z, pack, v, f, eps, sign)
    assert torch.equal(out2, out)
    assert float((logdet2 - logdet).abs().max()) <= 1e-6 * max(1.0, float(logdet.abs().max()))
    out3, logdet3 = ops.su3_heads_vupdate(z, pack, v, f, eps, sign)
    assert torch.equal(out3, out2) and torch.equal(logdet3, logdet2), 'run-to-run bit-reproducible'


@pytest.mark.parametrize('nb,xdim,hidden', [(64, 1152, 256), (70, 4320, 128), (5, 256, 64)])
@pytest.mark.parametrize('sign1,sign2,negate', [(+1, +1, False), (+1, -1, True), (-1, -1, False)])
def test_paired_heads_update_equals_two_single_updates(nb, xdim, hidden, sign1, sign2, negate):
    """l2b_su3_heads_vupdate_pair: two consecutive momentum updates on the same (s, t, q, F) in one pass -- the
    pairs of an L2HMC sweep (dynamics.py:1187-1228) incl. the v -> -v of the turn-around (dynamics.py:1002)"""
    from l2hmc_b200 import ops
    w, b, cs, cq, z, v, f = _case(nb, xdim, hidden, seed=3 * nb + hidden)
    pack = ops.vnet_pack_heads(w[0], w[1], w[2], b[0], b[1], b[2], cs, cq, 0.7, 1.3, 0.9)
    e1, e2 = 0.11, torch.tensor(0.07, dtype=torch.float64, device=DEV)      # by value and device-resident
    v1, ld1 = ops.su3_heads_vupdate(z, pack, v, f, e1, sign1)
    if negate:
        v1 = -v1
    v2, ld2 = ops.su3_heads_vupdate(z, pack, v1, f, e2, sign2)
    vp, ldp = ops.su3_heads_vupdate_pair(z, pack, v, f, e1, sign1, e2, sign2, negate_between=negate)
    assert torch.equal(vp, v2), 'same arithmetic per update'
    assert float((ldp - (ld1 + ld2)).abs().max()) <= 2e-6 * max(1.0, float((ld1 + ld2).abs().max()))


def test_paired_link_update_equals_two_masked_updates():
    """l2b_su3_update_gauge_planar_pair == l2b_su3_update_gauge_planar(mask) then (1 - mask), both orders"""
    from l2hmc_b200 import ops
    torch.manual_seed(12)
    nb, shape = 3, [4, 2, 4, 6]
    dev_ = torch.device(DEV)
    x = ops.su3_project(torch.complex(torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev_),
                                      torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev_)))
    p = ops.su3_rand_momentum(nb, shape, 9, 0, dev_) * torch.rand(nb, 4, *shape, 3, 3, device=dev_, dtype=torch.float64)
    xs, ps = ops.su3_aos_to_soa(x), ops.su3_aos_to_soa(p)
    mask = (torch.rand(xs[0].numel(), device=dev_) > 0.5).float()
    eps = torch.tensor(0.13, dtype=torch.float64, device=dev_)
    for first_c in (False, True):
        for sign in (+1.0, -1.0):
            a = ops.su3_update_gauge_planar(xs, ps, eps, mask, first_c, eps_mult=sign)
            a = ops.su3_update_gauge_planar(a, ps, eps, mask, not first_c, eps_mult=sign)
            b = ops.su3_update_gauge_planar_pair(xs, ps, eps, mask, first_c, eps_mult=sign)
            assert torch.equal(a, b)


def test_heads_unsupported_hidden_raises():
    from l2hmc_b200 import ops
    w, b, cs, cq, z, v, f = _case(4, 128, 260, seed=1)
    pack = ops.vnet_pack_heads(w[0], w[1], w[2], b[0], b[1], b[2], cs, cq)
    with pytest.raises(ops.L2BError):
        ops.su3_heads_vupdate(z, pack, v, f, 0.1, 1)


def _vnet(xshape, units, seed):
    from l2hmc_b200.configs import NetworkConfig, NetWeight
    from l2hmc_b200.network.pytorch.network import LeapfrogLayer
    torch.manual_seed(seed)
    net = LeapfrogLayer(xshape=xshape, network_config=NetworkConfig(units=list(units), activation_fn='tanh',
                                                                    dropout_prob=0.0, use_batch_norm=False),
                        net_weight=NetWeight(0.8, 1.1, 0.9)).to(DEV)
    V = int(np.prod(xshape[2:6]))
    with torch.no_grad():
        net((torch.zeros(2, 4 * V * 8, device=DEV), torch.zeros(2, 4 * V * 8, device=DEV)))   # materialise lazy layers
        for p in net.parameters():                      # bf16-representable parameters: the fused and the
            p.copy_(p.to(torch.bfloat16).to(p.dtype))   # unfused path then see identical operands
        net.scale.coeff.normal_(0, 0.1)
        net.transf.coeff.normal_(0, 0.1)
    return net


@pytest.mark.parametrize('sign', [+1, -1])
def test_fused_heads_function_matches_unfused_path_and_gradients(sign):
    """SU3HeadsVUpdate (tcgen05 heads + epilogue, recompute backward) against
    LeapfrogLayer.heads + SU3VUpdate on bf16-representable operands: values and every gradient"""
    from l2hmc_b200 import autograd as ag
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        nb, shape, hidden = 6, (2, 4, 2, 3), 64
        xshape = (nb, 4, *shape, 3, 3)
        xdim = 4 * int(np.prod(shape)) * 9
        net = _vnet(xshape, (hidden,), seed=5)
        g = torch.Generator(device='cpu').manual_seed(9)
        z0 = torch.tanh(torch.randn(nb, hidden, generator=g)).to(torch.bfloat16).float().to(DEV)
        v0 = torch.complex(torch.randn(nb, xdim, generator=g, dtype=torch.float64),
                           torch.randn(nb, xdim, generator=g, dtype=torch.float64)).to(DEV)
        f0 = torch.complex(torch.randn(nb, xdim, generator=g, dtype=torch.float64),
                           torch.randn(nb, xdim, generator=g, dtype=torch.float64)).to(DEV)
        wv = torch.complex(torch.randn(nb, xdim, generator=g, dtype=torch.float64),
                           torch.randn(nb, xdim, generator=g, dtype=torch.float64)).to(DEV)
        wl = torch.randn(nb, generator=g, dtype=torch.float64).to(DEV)
        params = [p for p in net.head_params()]

        def run(fused):
            z = z0.clone().requires_grad_(True)
            v = v0.clone().requires_grad_(True)
            f = f0.clone().requires_grad_(True)
            eps = torch.tensor(0.13, dtype=torch.float64, device=DEV, requires_grad=True)
            if fused:      # fields in their lattice shape, as Dynamics passes them (the adjoint kernel needs it)
                out, ld = ag.SU3HeadsVUpdate.apply(z, v.reshape(nb, 4, *shape, 3, 3), f.reshape(nb, 4, *shape, 3, 3),
                                                   eps, sign, None, net, *params)
                out = out.reshape(nb, xdim)
            else:
                s, t, q = net.heads(z)
                out, ld = ag.SU3VUpdate.apply(v.reshape(nb, 4, *shape, 3, 3), f.reshape(nb, 4, *shape, 3, 3), s, t, q,
                                              eps, sign)
                out = out.reshape(nb, xdim)
            loss = (out * wv.conj()).real.sum() + (ld * wl).sum()
            grads = torch.autograd.grad(loss, [z, v, f, eps] + params)
            return out.detach(), ld.detach(), grads
        o1, l1, g1 = run(True)
        o0, l0, g0 = run(False)
        assert float((o1 - o0).abs().max()) < 1e-5 * max(1.0, float(o0.abs().max()))
        assert float((l1 - l0).abs().max()) < 1e-5 * max(1.0, float(l0.abs().max()))
        for a, b in zip(g1, g0):
            assert a.shape == b.shape and a.dtype == b.dtype
            assert float((a - b).abs().max()) < 2e-5 * max(1.0, float(b.abs().max()))
    finally:
        torch.set_default_dtype(old)


def test_dynamics_uses_tensor_core_heads_under_bf16_autocast():
    """BASELINE cfg 5 path: under bf16 autocast the SU(3) v-update runs k_heads_vupdate; the
    sweep agrees with the unfused bf16 path to bf16 accuracy and the train step yields finite grads"""
    from tests._helpers import _su3_trainer
    from l2hmc_b200 import _lib
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        torch.manual_seed(1)
        np.random.seed(1)
        tr, lat = _su3_trainer(nb=4, units=(32,), autocast=torch.bfloat16)
        dyn = tr.dynamics
        x = lat.random().to(torch.complex128)
        v = lat.random_momentum()
        beta = torch.tensor(6.0)
        from l2hmc_b200.dynamics.pytorch.dynamics import State
        outs = {}
        for mode in ('never', 'auto'):
            dyn.tensor_core_heads = mode
            n0 = _lib.launch_count()
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                st, met = dyn.transition_kernel_fb(State(x, v, beta))
            outs[mode] = (st.x, st.v, met['sumlogdet'], _lib.launch_count() - n0)
        assert float((outs['auto'][1] - outs['never'][1]).abs().max()) < 5e-2 * float(outs['never'][1].abs().max())
        assert float((outs['auto'][0] - outs['never'][0]).abs().max()) < 5e-2
        assert outs['auto'][3] != outs['never'][3], 'the fused kernel replaces k_vupdate launches'
        dyn.tensor_core_heads = 'auto'
        xo, m = tr.train_step((x, beta))
        assert torch.isfinite(m['loss'])
        gr = {n: p.grad for n, p in dyn.named_parameters() if p.grad is not None}
        assert any('vnet.scale.layer.weight' in n for n in gr) and any('veps' in n for n in gr)
        assert all(torch.isfinite(t).all() for t in gr.values())
    finally:
        torch.set_default_dtype(old)


def test_planar_inference_sweep_matches_boundary_layout_sweep():
    """Dynamics.transition_kernel_fb with the state kept planar (no layout conversion around the stencil
    kernels, head weights packed with permuted rows) == the boundary-layout sweep, link for link"""
    from tests._helpers import _su3_trainer
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    from l2hmc_b200 import _lib
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        torch.manual_seed(4)
        np.random.seed(4)
        tr, lat = _su3_trainer(nb=3, shape=(4, 2, 4, 6), units=(32,), nlf=2, autocast=torch.bfloat16)
        dyn = tr.dynamics
        dyn.eval()
        x = lat.random().to(torch.complex128)
        v = lat.random_momentum()
        beta = torch.tensor(5.9)
        res = {}
        # (planar sweep, paired updates, tensor-core input layer)
        settings = {'boundary': ('never', 'never', 'never'), 'planar': ('auto', 'never', 'never'),
                    'paired': ('auto', 'auto', 'never'), 'fused': ('auto', 'auto', 'auto')}
        for name, (planar, pair, tci) in settings.items():
            dyn.planar_sweep, dyn.pair_updates, dyn.tensor_core_input = planar, pair, tci
            n0 = _lib.launch_count()
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                st, met = dyn.transition_kernel_fb(State(x, v, beta))
            res[name] = (st.x.reshape(x.shape), st.v.reshape(x.shape), met['sumlogdet'], met['acc'],
                         _lib.launch_count() - n0)
        dyn.planar_sweep = dyn.pair_updates = dyn.tensor_core_input = 'auto'
        a, b, c, d = res['boundary'], res['planar'], res['paired'], res['fused']
        assert b[4] < a[4], 'fewer launches: no conversions around the force'
        assert c[4] < b[4], 'fewer launches: two updates per pass'
        assert float((a[0] - b[0]).abs().max()) < 1e-13 and float((a[1] - b[1]).abs().max()) < 1e-12
        assert float((a[2] - b[2]).abs().max()) < 1e-5 * max(1.0, float(a[2].abs().max()))
        assert float((a[3] - b[3]).abs().max()) < 1e-5
        # paired passes: the same arithmetic per update on the same network outputs -> the same links and momenta
        assert float((b[0] - c[0]).abs().max()) < 1e-13 and float((b[1] - c[1]).abs().max()) < 1e-12
        assert float((b[2] - c[2]).abs().max()) < 1e-5 * max(1.0, float(b[2].abs().max()))
        # tensor-core input layer: another bf16 GEMM (fp32 accumulation in another order, z rounded to bf16 either
        # way) -> agreement at bf16 level only; the kernel itself is pinned by test_input_layer_matches_float64_reference
        assert float((c[0] - d[0]).abs().max()) < 2e-2 and float((c[1] - d[1]).abs().max()) < 5e-2 * float(c[1].abs().max())
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('nb,shape,hidden,act', [(256, (8, 8, 8, 8), 256, 'tanh'), (32, (8, 8, 8, 8), 256, 'tanh'),
                                                  (3, (4, 2, 4, 6), 32, 'relu'), (20, (4, 4, 4, 4), 136, 'swish'),
                                                  (17, (2, 2, 2, 4), 8, 'leaky_relu')])
def test_input_layer_matches_float64_reference(nb, shape, hidden, act):
    """tcgen05 split-K input layer (l2b_su3_input_layer fed by l2b_su3_project_vec_planar_lm) against a float64
    evaluation of  act(W_x vec_x + b_x + W_v vec_f + b_v)  on the SAME bf16-rounded operands: what is checked is the
    kernel (operand images, UMMA descriptors, split-K partials, reduction, epilogue), not bf16 rounding.
    Reference semantics: network/pytorch/network.py:349-451 with the inputs of dynamics.py:1142-1160."""
    from l2hmc_b200 import ops
    torch.manual_seed(hidden + nb)
    V = int(np.prod(shape))
    K = 8 * 4 * V
    dev_ = torch.device(DEV)
    x = ops.su3_project(torch.complex(torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev_),
                                      torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev_)))
    f = ops.su3_rand_momentum(nb, list(shape), 5, 0, dev_)
    xs, fs = ops.su3_aos_to_soa(x), ops.su3_aos_to_soa(f)
    wx = (torch.randn(hidden, K, device=dev_) / K ** 0.5)
    wv = (torch.randn(hidden, K, device=dev_) / K ** 0.5)
    bx, bv = torch.randn(hidden, device=dev_), torch.randn(hidden, device=dev_)
    pack = ops.su3_input_pack(wx, wv, bx, bv, act)
    ax, af = ops.su3_project_vec_planar_lm(xs), ops.su3_project_vec_planar_lm(fs)
    nbp = ops.input_nb_pad(nb)
    assert ax.shape == (4 * V, nbp, 8) and ax.dtype == torch.bfloat16
    # the link-major image holds exactly the vec8 the unfused path produces, transposed
    vx = ops.su3_project_vec_planar(xs, torch.bfloat16).reshape(nb, 4 * V, 8)
    vf = ops.su3_project_vec_planar(fs, torch.bfloat16).reshape(nb, 4 * V, 8)
    assert torch.equal(ax[:, :nb].permute(1, 0, 2), vx) and torch.equal(af[:, :nb].permute(1, 0, 2), vf)
    assert float(ax[:, nb:].abs().sum()) == 0.0
    z = ops.su3_input_layer(ax, af, pack, nb)
    assert z.shape == (nb, hidden) and z.dtype == torch.bfloat16
    r = lambda t: t.to(torch.bfloat16).double()   # noqa: E731
    pre = r(vx).reshape(nb, K) @ r(wx).t() + r(vf).reshape(nb, K) @ r(wv).t() + bx.double() + bv.double()
    fn = {'tanh': torch.tanh, 'relu': torch.relu, 'swish': torch.nn.functional.silu,
          'leaky_relu': torch.nn.functional.leaky_relu}[act]
    want = fn(pre)
    # fp32 accumulation over K = 2 * 8 * 4 V terms, then one bf16 rounding of z (2^-9 relative)
    err = (z.double() - want).abs()
    assert float(err.max()) < 6e-3 * max(1.0, float(want.abs().max())), float(err.max())
    z2 = ops.su3_input_layer(ax, af, pack, nb)
    assert torch.equal(z, z2), 'fixed-order split-K reduction: bit-reproducible'


def test_bf16_vnet_inputs_of_unitary_links_equal_the_full_projection():
    """bf16 producers of the vnet inputs return links that are already in SU(3) as they are (project_su<NEAR>):
    against the float64 full projection (dynamics.py:1154-1156, utils.py:227-346) rounded to bf16 the result may
    differ only where the ~1e-8 noise of the closed-form eigen-solve at a triply degenerate X^+X crosses a bf16
    rounding boundary; non-unitary inputs (the force) still take the full path"""
    from l2hmc_b200 import ops
    from oracle import su3 as osu3
    rng = np.random.default_rng(11)
    nb, shape = 4, (4, 4, 2, 6)
    x = torch.from_numpy(osu3.random_su3(rng, (nb, 4, *shape, 3, 3))).to(DEV)
    f = torch.from_numpy(osu3.random_momentum(rng, (nb, 4, *shape, 3, 3))).to(DEV)
    for field, unitary in ((x, True), (f, False)):
        want = ops.su3_project_vec(field, torch.float64)
        got = ops.su3_project_vec(field, torch.bfloat16)
        ulp = want.abs().clamp(min=2.0 ** -120) * 2.0 ** -8            # one bf16 ulp at each element, at least
        assert bool(((got.double() - want).abs() <= ulp).all())
        exact = got == want.to(torch.bfloat16)
        assert float(exact.float().mean()) > 0.999
        if unitary:                                                   # and the vec8 of the link itself, bit for bit
            direct = ops.su3_to_vec(field).to(torch.bfloat16) if hasattr(ops, 'su3_to_vec') else None
            if direct is not None:
                assert torch.equal(got, direct.reshape(got.shape))
        xs = ops.su3_aos_to_soa(field)
        V = int(np.prod(shape))
        pl = ops.su3_project_vec_planar(xs, torch.bfloat16).reshape(nb, 4 * V, 8)
        assert torch.equal(pl.reshape(-1), got.reshape(-1))
        lm = ops.su3_project_vec_planar_lm(xs)                        # [link][chain_pad][8]
        assert torch.equal(lm[:, :nb].permute(1, 0, 2).reshape(-1), got.reshape(-1))
