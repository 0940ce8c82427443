"""oracle/make_golden.py -- TEST INFRASTRUCTURE.

Runs the reference's OWN PyTorch hot path (via `oracle/ref_shim.py`, i.e. the
unmodified modules copied to `oracle/_ref`) on small seeded inputs and freezes
inputs + outputs into `tests/golden/*.npz`.  The reference ships no tests or
golden vectors, so these files are what pins parity:

  * `tests/test_oracle_golden.py`            numpy oracle == golden            (CPU)
  * `tests/test_hostemu.py`                  kernel bodies on the host == golden (CPU)
  * `tests/test_host_logic_cpu_emulated.py`  our Dynamics / Trainer on CPU stand-ins == golden (CPU)
  * `tests/test_gpu_*.py`                    CUDA path == golden                 (-m gpu)

Contents: lattice / group functions, HMC trajectories, the merged and un-merged L2HMC kernels, verbose per-step
histories, LatticeLoss, parameter / input gradients from the reference's autograd, per-function VJPs, and the
improved action (c1 != 0) incl. an HMC case whose acceptance is not trivially 1.  Seeds are fixed: regenerating
reproduces every existing key bit for bit (checked whenever keys are added).

Run here (needs /root/reference):   python oracle/make_golden.py
It re-executes itself once per default dtype, because the reference freezes
dtype-typed module constants at import (`group/su3/pytorch/utils.py:28-36`).
"""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
GOLD = ROOT / 'tests' / 'golden'
sys.path.insert(0, str(ROOT))


def _np(t):
    return t.detach().cpu().numpy()


def _sd(dyn) -> dict:
    """reference state_dict without the duplicated `networks.` prefix"""
    return {f'sd/{k}': _np(v) for k, v in dyn.state_dict().items()
            if not k.startswith('networks.')}


def _randomise_coeffs(dyn, torch):
    """`coeff` of ScaledTanh initialises to 0 and BN running stats to (0, 1);
    perturb them so the goldens exercise those terms."""
    with torch.no_grad():
        for n, p in dyn.named_parameters():
            if n.endswith('coeff'):
                p.normal_(0, 0.1)
        for n, b in dyn.named_buffers():
            if 'running_mean' in n:
                b.normal_(0, 0.1)
            if 'running_var' in n:
                b.uniform_(0.5, 1.5)


# ---------------------------------------------------------------------------
def su3_goldens(ref, torch):
    assert torch.get_default_dtype() == torch.float64
    shape, nb, beta = [4, 3, 2, 5], 2, 5.7
    lat = ref.LatticeSU3(nb, shape)
    g = lat.g
    torch.manual_seed(20261017)
    x = lat.random()
    v = lat.random_momentum()
    y = torch.complex(torch.randn(*lat._shape), torch.randn(*lat._shape))
    b = torch.tensor(beta)
    f = lat.grad_action(x, b).detach()
    x = x.detach()
    ca, cm = g.checkSU(y)
    out = dict(
        shape=np.array(shape), beta=beta, x=_np(x), v=_np(v), y=_np(y),
        wloops=_np(lat.wilson_loops(x)), action=_np(lat.action(x, b)),
        plaqs=_np(lat._plaquettes(x)), intQ=_np(lat.int_charges(x)),
        sinQ=_np(lat.sin_charges(x)), ke=_np(g.kinetic_energy(v)),
        force=_np(f), upd=_np(g.update_gauge(x, 0.1 * v)),
        expv=_np(g.exp(0.25 * v)), expy=_np(g.exp(y)),
        projsu_y=_np(g.projectSU(y)), tah_y=_np(g.projectTAH(y)),
        vec_x=_np(g.group_to_vec(x)), vec_f=_np(g.group_to_vec(f)),
        vec2su3=_np(ref.su3utils.vec_to_su3(g.group_to_vec(x))),
        checksu_avg=_np(ca), checksu_max=_np(cm),
    )
    # cold start known answers: S = -6 beta V, F = 0
    cold = torch.eye(3, dtype=torch.complex128).expand(*lat._shape).contiguous()
    out['cold_action'] = _np(lat.action(cold, b))
    out['cold_force_max'] = float(lat.grad_action(cold.clone(), b).abs().max())
    # plain HMC trajectory with explicit momenta
    cfg = ref.DynamicsConfig(
        nchains=nb, group='SU3', latvolume=shape, nleapfrog=2, eps=0.05,
        eps_hmc=0.1, verbose=False, use_split_xnets=False,
        use_separate_networks=False, merge_directions=True)
    dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
    st = ref.State(x=x, v=v, beta=b)
    for nlf, eps in ((1, 0.1), (4, 0.05)):
        sp, met = dyn.transition_kernel_hmc(st, eps=eps, nleapfrog=nlf)
        k = f'hmc{nlf}'
        out[f'{k}_eps'] = eps
        out[f'{k}_x'] = _np(sp.x)
        out[f'{k}_v'] = _np(sp.v)
        out[f'{k}_acc'] = _np(met['acc'])
        out[f'{k}_h0'] = _np(dyn.hamiltonian(st))
        out[f'{k}_h1'] = _np(dyn.hamiltonian(sp))
    # warm start (plaquette ~0.8 > equilibrium) so that 0 < acc < 1 is exercised
    cold = torch.eye(3, dtype=torch.complex128).expand(*lat._shape).contiguous()
    xw = g.update_gauge(cold, 0.2 * lat.random_momentum())
    vw = lat.random_momentum()
    out['xw'], out['vw'] = _np(xw), _np(vw)
    stw = ref.State(x=xw, v=vw, beta=b)
    spw, metw = dyn.transition_kernel_hmc(stw, eps=0.05, nleapfrog=4)
    out['hmcw_eps'], out['hmcw_nlf'] = 0.05, 4
    out['hmcw_x'], out['hmcw_v'] = _np(spw.x), _np(spw.v)
    out['hmcw_acc'] = _np(metw['acc'])
    out['hmcw_h0'], out['hmcw_h1'] = _np(dyn.hamiltonian(stw)), _np(dyn.hamiltonian(spw))
    np.savez_compressed(GOLD / 'su3_f64.npz', **out)

    # ---- L2HMC forward sweep on a smaller lattice, vnet units [8] ---------
    shape, nb, nlf = [2, 4, 2, 3], 2, 2
    lat = ref.LatticeSU3(nb, shape)
    cfg = ref.DynamicsConfig(
        nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=0.05,
        eps_hmc=0.1, verbose=False, use_split_xnets=False,
        use_separate_networks=False, merge_directions=True)
    V = int(np.prod(shape))
    ispec = ref.InputSpec(xshape=cfg.xshape,
                          xnet={'x': [4 * V * 8], 'v': [4 * V * 8]},
                          vnet={'x': [4 * V * 8], 'v': [4 * V * 8]})
    ncfg = ref.NetworkConfig(units=[8], activation_fn='tanh',
                             dropout_prob=0.0, use_batch_norm=False)
    nw = ref.NetWeights(x=ref.NetWeight(0., 1., 1.), v=ref.NetWeight(1., 1., 1.))
    torch.manual_seed(7)
    np.random.seed(7)
    fac = ref.NetworkFactory(input_spec=ispec, network_config=ncfg,
                             conv_config=None, net_weights=nw)
    dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    _randomise_coeffs(dyn, torch)
    dyn.eval()
    x = lat.random().detach()
    v = lat.random_momentum()
    b = torch.tensor(6.0)
    st = ref.State(x=x, v=v, beta=b)
    out = dict(shape=np.array(shape), beta=6.0, nlf=nlf, x=_np(x), v=_np(v),
               masks=np.stack([_np(m) for m in dyn.masks]),
               xeps=np.array([float(e) for e in dyn.xeps]),
               veps=np.array([float(e) for e in dyn.veps]),
               nw_x=np.array([0., 1., 1.]), nw_v=np.array([1., 1., 1.]))
    # the SU(3) x-update never calls xnet (dynamics.py:1420-1425): skip its weights
    out.update({k: a for k, a in _sd(dyn).items() if not k.startswith('sd/xnet')})
    s1, ld1 = dyn._update_v_fwd(0, st)
    out['vfwd_v'], out['vfwd_logdet'] = _np(s1.v), _np(ld1)
    s1b, ld1b = dyn._update_v_bwd(1, st)
    out['vbwd_v'], out['vbwd_logdet'] = _np(s1b.v), _np(ld1b)
    m, _ = dyn._get_mask(0)
    s2, _ = dyn._update_x_fwd(0, st, m, first=True)
    out['xfwd_x'] = _np(s2.x)
    s2b, _ = dyn._update_x_bwd(0, st, m, first=True)
    out['xbwd_x'] = _np(s2b.x)
    sp, met = dyn.transition_kernel_fb(st)
    out['fb_x'], out['fb_v'] = _np(sp.x), _np(sp.v)
    out['fb_acc'], out['fb_sumlogdet'] = _np(met['acc']), _np(met['sumlogdet'])
    # gradients through the reference's own autograd graph of a scalar touching
    # x_prop, acc (both Hamiltonians), sumlogdet and the plaquette sums
    wgt = torch.randn(*x.shape)
    out['loss_w'] = _np(wgt)
    loss = ((met['acc'] * (sp.x.real * wgt).flatten(1).sum(1)).sum() + met['sumlogdet'].sum()
            + 0.01 * (met['acc'] * lat.wilson_loops(sp.x).real.sum((0, 2, 3, 4, 5))).sum())
    named = [(n[len('networks.'):] if n.startswith('networks.') else n, p)
             for n, p in dyn.named_parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, [p for _, p in named] + [x], allow_unused=True, retain_graph=True)
    out['loss'] = _np(loss)
    for (n, _), g_ in zip(named, grads[:-1]):
        if g_ is not None:
            out['grad/' + n] = _np(g_)
    out['grad_x'] = _np(grads[-1])
    if ref.LatticeLoss is not None:      # conf/loss/su3.yaml: plaq 0.1 + rmse 0.1
        lcfg = ref.cfgs.LossConfig(use_mixed_loss=False, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1)
        out['lattice_loss'] = _np(ref.LatticeLoss(lat, lcfg)(x_init=x, x_prop=sp.x, acc=met['acc']))
    np.savez_compressed(GOLD / 'su3_l2hmc_f64.npz', **out)


def su3_adjoint_goldens(ref, torch):
    """vector-Jacobian products the reference's autograd gives for the per-link maps
    projectSU / group_to_vec (utils.py:227-346,394-420) and for `wilson_loops`
    (lattice.py:157-199): what the hand-written adjoint kernels must reproduce."""
    assert torch.get_default_dtype() == torch.float64
    torch.manual_seed(424242)
    g = ref.LatticeSU3(1, [2, 2, 2, 2]).g
    n = 48
    xg = torch.complex(torch.randn(n, 3, 3), torch.randn(n, 3, 3))               # generic
    xs = g.projectSU(torch.complex(torch.randn(n, 3, 3), torch.randn(n, 3, 3)))
    xs = xs + 1e-3 * torch.complex(torch.randn(n, 3, 3), torch.randn(n, 3, 3))   # near SU(3)
    xa = g.projectTAH(torch.complex(torch.randn(n, 3, 3), torch.randn(n, 3, 3)))  # a force
    x = torch.cat([xg, xs, xa]).detach().requires_grad_(True)
    gmat = torch.complex(torch.randn(3 * n, 3, 3), torch.randn(3 * n, 3, 3))
    gvec = torch.randn(3 * n, 8)
    gx_mat, = torch.autograd.grad(g.projectSU(x), x, grad_outputs=gmat)
    gx_vec, = torch.autograd.grad(g.group_to_vec(x), x, grad_outputs=gvec)
    out = dict(x=_np(x), gmat=_np(gmat), gvec=_np(gvec), gx_mat=_np(gx_mat), gx_vec=_np(gx_vec))
    shape, nb = [4, 3, 2, 5], 2
    lat = ref.LatticeSU3(nb, shape)
    u = lat.random().detach().requires_grad_(True)
    w = lat.wilson_loops(u)
    gw = torch.complex(torch.randn(*w.shape), torch.randn(*w.shape))
    gu, = torch.autograd.grad(w, u, grad_outputs=gw)
    out.update(wl_shape=np.array(shape), wl_x=_np(u), wl_gw=_np(gw), wl_gx=_np(gu))
    np.savez_compressed(GOLD / 'su3_adjoint_f64.npz', **out)
    # ---- rectangle (c1 != 0, DBW2 value) action: loops, action, force, one HMC trajectory ----
    shape, nb, beta, c1 = [2, 4, 3, 2], 2, 5.7, -0.331
    lat = ref.LatticeSU3(nb, shape, c1=c1)
    xr = lat.random().detach()
    vr = lat.random_momentum()
    b = torch.tensor(beta)
    ps, rs = lat._wilson_loops(xr, needs_rect=True)
    cfg = ref.DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=2, eps=0.05, eps_hmc=0.05,
                             verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
    sp, met = dyn.transition_kernel_hmc(ref.State(x=xr, v=vr, beta=b), eps=0.05, nleapfrog=3)
    # a second trajectory from a smooth start (links exp(0.2 P) near the identity) with a short trajectory, so that
    # dH > 0 and `acc` is not trivially 1 (a hot start only relaxes: dH << 0).  NB the reference's Dynamics
    # integrates with the force of ITS OWN lattice (c1 = 0, dynamics.py:134,1499) and takes the energies from
    # `potential_fn` (c1 != 0)
    x2 = torch.linalg.matrix_exp(0.2 * lat.random_momentum()).detach()
    v2 = lat.random_momentum()
    st2 = ref.State(x=x2, v=v2, beta=b)
    sp2, met2 = dyn.transition_kernel_hmc(st2, eps=0.01, nleapfrog=3)
    np.savez_compressed(GOLD / 'su3_c1_f64.npz', shape=np.array(shape), beta=beta, c1=c1, x=_np(xr), v=_np(vr),
                        rects=_np(rs), action=_np(lat.action(xr, b)), force=_np(lat.grad_action(xr.clone(), b)),
                        hmc_x=_np(sp.x), hmc_v=_np(sp.v), hmc_acc=_np(met['acc']),
                        hmc2_x0=_np(x2), hmc2_v0=_np(v2), hmc2_x=_np(sp2.x), hmc2_v=_np(sp2.v),
                        hmc2_acc=_np(met2['acc']), hmc2_h0=_np(dyn.hamiltonian(st2)),
                        hmc2_h1=_np(dyn.hamiltonian(ref.State(x=sp2.x.reshape(xr.shape), v=sp2.v, beta=b))))


def u1_goldens(ref, torch, tag: str):
    shape, nb, beta, nlf = [8, 6], 3, 4.0, 2
    lat = ref.LatticeU1(nb, shape)
    torch.manual_seed(11)
    x = lat.random()
    b = torch.tensor(beta)
    out = dict(
        shape=np.array(shape), beta=beta, x=_np(x),
        wloops=_np(lat.wilson_loops(x)), action=_np(lat.action(x, b)),
        force=_np(lat.grad_action(x.clone(), b)), plaqs=_np(lat.plaqs(x=x)),
        sinQ=_np(lat.sin_charges(x=x)), intQ=_np(lat.int_charges(x=x)),
        wloops4x4=_np(lat.wilson_loops4x4(x)),
        compat=_np(lat.g.compat_proj(3.0 * x)),
    )
    convs = {'dense': None,
             'conv': dict(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2])}
    for name, conv in convs.items():
        cfg = ref.DynamicsConfig(
            nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1,
            eps_hmc=0.1, use_ncp=True, verbose=False, use_split_xnets=True,
            use_separate_networks=True, merge_directions=True)
        xdim = cfg.xdim
        ispec = ref.InputSpec(xshape=cfg.xshape,
                              xnet={'x': [xdim, 2], 'v': [xdim]},
                              vnet={'x': [xdim], 'v': [xdim]})
        ncfg = ref.NetworkConfig(units=[16, 12], activation_fn='leaky_relu',
                                 dropout_prob=0.2, use_batch_norm=True)
        ccfg = ref.ConvolutionConfig(**conv) if conv else None
        torch.manual_seed(3)
        np.random.seed(3)
        fac = ref.NetworkFactory(input_spec=ispec, network_config=ncfg,
                                 conv_config=ccfg, net_weights=None)
        dyn = ref.Dynamics(potential_fn=lat.action, config=cfg,
                           network_factory=fac)
        _randomise_coeffs(dyn, torch)
        dyn.eval()
        v = lat.g.random_momentum(list(cfg.xshape))
        st = ref.State(x=x, v=v, beta=b)
        if name == 'dense':
            out['v'] = _np(v)
            sp, met = dyn.transition_kernel_hmc(st, eps=0.1, nleapfrog=5)
            out['hmc_x'], out['hmc_v'] = _np(sp.x), _np(sp.v)
            out['hmc_acc'] = _np(met['acc'])
            out['hmc_h0'] = _np(dyn.hamiltonian(st))
            out['hmc_h1'] = _np(dyn.hamiltonian(
                ref.State(sp.x.reshape(x.shape), sp.v, b)))
        pre = f'{name}/'
        out[pre + 'v'] = _np(v)
        out[pre + 'masks'] = np.stack([_np(m) for m in dyn.masks])
        out[pre + 'xeps'] = np.array([float(e) for e in dyn.xeps])
        out[pre + 'veps'] = np.array([float(e) for e in dyn.veps])
        for k, val in _sd(dyn).items():
            out[pre + k] = val
        sp, met = dyn.transition_kernel_fb(st)
        out[pre + 'fb_x'], out[pre + 'fb_v'] = _np(sp.x), _np(sp.v)
        out[pre + 'fb_acc'] = _np(met['acc'])
        out[pre + 'fb_sumlogdet'] = _np(met['sumlogdet'])
        m, _ = dyn._get_mask(0)
        s2, ld2 = dyn._update_x_fwd(0, st, m, first=True)
        out[pre + 'xfwd_x'], out[pre + 'xfwd_logdet'] = _np(s2.x), _np(ld2)
        s3, ld3 = dyn._update_x_bwd(1, st, m, first=False)
        out[pre + 'xbwd_x'], out[pre + 'xbwd_logdet'] = _np(s3.x), _np(ld3)
        s4, ld4 = dyn._update_v_fwd(0, st)
        out[pre + 'vfwd_v'], out[pre + 'vfwd_logdet'] = _np(s4.v), _np(ld4)
        # gradients of a scalar that touches every output of the sweep (x_prop, acc,
        # sumlogdet, wilson loops), through the reference's own autograd graph
        sp, met = dyn.transition_kernel_fb(st)
        xp = sp.x.flatten(1)
        loss = ((met['acc'] * xp.cos().sum(1)).sum() + met['sumlogdet'].sum()
                + (met['acc'] * lat.wilson_loops(sp.x).sin().sum((1, 2))).sum())
        # named_parameters() de-duplicates the twice-registered nets and keeps the `networks.` alias
        named = [(n[len('networks.'):] if n.startswith('networks.') else n, p)
                 for n, p in dyn.named_parameters() if p.requires_grad]
        grads = torch.autograd.grad(loss, [p for _, p in named] + [x], allow_unused=True)
        out[pre + 'loss'] = _np(loss)
        for (n, _), g_ in zip(named, grads[:-1]):
            if g_ is not None:
                out[pre + 'grad/' + n] = _np(g_)
        out[pre + 'grad_x'] = _np(grads[-1])
        if ref.LatticeLoss is not None:  # conf/loss/default.yaml: mixed loss, charge_weight 0.01
            lcfg = ref.cfgs.LossConfig(use_mixed_loss=True, charge_weight=0.01, rmse_weight=0.0, plaq_weight=0.0)
            out[pre + 'lattice_loss'] = _np(ref.LatticeLoss(lat, lcfg)(x_init=x, x_prop=sp.x, acc=met['acc']))
            lcfg2 = ref.cfgs.LossConfig(use_mixed_loss=False, charge_weight=0.05, rmse_weight=0.0, plaq_weight=0.0)
            out[pre + 'lattice_loss2'] = _np(ref.LatticeLoss(lat, lcfg2)(x_init=x, x_prop=sp.x, acc=met['acc']))
        if name == 'dense':
            # the un-merged kernel (merge_directions = False: dynamics.py:1031-1063, note the swapped states in
            # its accept probability) and the verbose per-step metrics (dynamics.py:865-898,956-1029)
            for key, fwd in (('tkf', True), ('tkb', False)):
                sk, mk = dyn.transition_kernel(st, forward=fwd)
                out[f'{pre}{key}_x'], out[f'{pre}{key}_v'] = _np(sk.x), _np(sk.v)
                out[f'{pre}{key}_acc'], out[f'{pre}{key}_sumlogdet'] = _np(mk['acc']), _np(mk['sumlogdet'])
            cfgv = ref.DynamicsConfig(
                nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, eps_hmc=0.1, use_ncp=True,
                verbose=True, use_split_xnets=True, use_separate_networks=True, merge_directions=True)
            dynv = ref.Dynamics(potential_fn=lat.action, config=cfgv, network_factory=fac)
            dynv.load_state_dict(dyn.state_dict())
            dynv.masks = dyn.masks
            dynv.eval()
            for key, (sv, hv) in (('vfb', dynv.transition_kernel_fb(st)),
                                  ('vtk', dynv.transition_kernel(st, forward=True)),
                                  ('vhmc', dynv.transition_kernel_hmc(st, eps=0.1, nleapfrog=3))):
                out[f'{pre}{key}_x'] = _np(sv.x)
                for hk, hval in hv.items():
                    if isinstance(hval, torch.Tensor):
                        out[f'{pre}{key}/{hk}'] = _np(hval)
                    elif isinstance(hval, list):          # xeps / veps: lists of 0-dim parameters
                        out[f'{pre}{key}/{hk}'] = np.array([float(e) for e in hval])
    np.savez_compressed(GOLD / f'u1_{tag}.npz', **out)


def main():
    import torch
    from oracle import make_ref
    from oracle.ref_shim import load_reference
    if len(sys.argv) < 2:
        make_ref.build()
        GOLD.mkdir(parents=True, exist_ok=True)
        for tag in ('f64', 'f32', 'adjoint'):
            subprocess.check_call([sys.executable, __file__, tag])
        for f in sorted(GOLD.glob('*.npz')):
            print(f'{f.name:24s} {f.stat().st_size / 1024:8.1f} KiB')
        return
    tag = sys.argv[1]
    torch.set_num_threads(4)
    if tag == 'adjoint':       # python oracle/make_golden.py adjoint  (adds one file, leaves the others)
        return su3_adjoint_goldens(load_reference(torch.float64), torch)
    ref = load_reference(torch.float64 if tag == 'f64' else torch.float32)
    if tag == 'f64':
        su3_goldens(ref, torch)
    u1_goldens(ref, torch, tag)


if __name__ == '__main__':
    main()
