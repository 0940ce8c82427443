"""CPU tier: host logic of the force / vnet-input reuse between consecutive v-updates
(`Dynamics._force`, `_vnet_vecs`; `reuse_force = 'always'`, the default).  The methods only need
`grad_potential` and `group_to_vec`, so they run here on a stand-in object with counting stubs."""
import types

import torch

from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State


def make(mode):
    calls = {'force': 0, 'vec': 0}
    fake = types.SimpleNamespace(reuse_force=mode, _fcache=None)

    def grad_potential(x, beta):
        calls['force'] += 1
        return x * 2.0

    def group_to_vec(x, dt):
        calls['vec'] += 1
        return x.to(dt)
    fake.grad_potential = grad_potential
    fake.group_to_vec = group_to_vec
    fake._reuse_force = lambda: Dynamics._reuse_force(fake)
    fake._force = lambda st: Dynamics._force(fake, st)
    fake._vnet_vecs = lambda st, f, dt: Dynamics._vnet_vecs(fake, st, f, dt)
    return fake, calls


def test_default_recomputes_every_time():
    fake, calls = make('never')
    st = State(torch.randn(2, 4), torch.randn(2, 4), torch.tensor(1.0))
    f1, f2 = fake._force(st), fake._force(st)
    assert calls['force'] == 2 and torch.equal(f1, f2)
    fake._vnet_vecs(st, f1, torch.float32)
    fake._vnet_vecs(st, f2, torch.float32)
    assert calls['vec'] == 4 and fake._fcache is None


def test_reuse_hits_only_for_the_same_unmodified_links():
    fake, calls = make('always')
    beta = torch.tensor(1.0)
    x = torch.randn(2, 4)
    st = State(x, torch.randn(2, 4), beta)
    f1 = fake._force(st)
    st2 = State(st.x, -st.v, st.beta)            # v-update / turn-around: same links object
    f2 = fake._force(st2)
    assert f2 is f1 and calls['force'] == 1
    a = fake._vnet_vecs(st, f1, torch.float32)
    b = fake._vnet_vecs(st2, f2, torch.float32)
    assert a[0] is b[0] and a[1] is b[1] and calls['vec'] == 2
    c = fake._vnet_vecs(st2, f2, torch.float64)  # another net dtype: its own pair
    assert calls['vec'] == 4 and c[0].dtype == torch.float64
    # an x-update produces a new tensor -> miss
    st3 = State(x + 0.0, st.v, beta)
    f3 = fake._force(st3)
    assert f3 is not f1 and calls['force'] == 2
    # in-place modification of the cached links -> version bump -> miss
    st3.x.add_(1.0)
    f4 = fake._force(st3)
    assert calls['force'] == 3 and torch.equal(f4, st3.x * 2.0)
    # a different beta object, or a different grad mode -> miss
    fake._force(State(st3.x, st.v, torch.tensor(1.0)))
    assert calls['force'] == 4
    with torch.no_grad():
        fake._force(State(st3.x, st.v, fake._fcache['beta']))
    assert calls['force'] == 5
    # a force that is not the cached one (caller computed its own) is never paired with cached vecs
    other = st3.x * 3.0
    fake._vnet_vecs(State(st3.x, st.v, beta), other, torch.float32)
    assert calls['vec'] == 6


def test_planar_sweep_schedule_with_fake_ops(monkeypatch):
    """`_transition_kernel_fb_planar` executed on the CPU with stand-in ops (deterministic torch
    functions that count their calls): the default schedule evaluates the force once per v-update
    (4 nlf) and projects twice per v-update; with reuse it is once per distinct x (2 nlf + 1), and
    both give the same state, acceptance and log-Jacobian."""
    import l2hmc_b200.dynamics.pytorch.dynamics as dmod
    nb, n, nlf = 3, 10, 2
    calls = {}

    def count(name):
        calls[name] = calls.get(name, 0) + 1

    class FakeOps:
        @staticmethod
        def su3_aos_to_soa(t):
            return t.clone()

        @staticmethod
        def su3_soa_to_aos(t):
            return t.clone()

        @staticmethod
        def su3_force_planar(xs, beta):
            count('force')
            return torch.sin(xs) * beta

        @staticmethod
        def su3_project_vec_planar(t, dt):
            count('project')
            return (t * 0.5).to(dt)

        @staticmethod
        def su3_heads_vupdate(z, pack, v, f, eps, sign):
            count('vupdate')
            return v + sign * eps * (f + z.to(v.dtype)), (z.double() * sign * eps).sum(1)

        @staticmethod
        def su3_update_gauge_planar(xs, vs, eps, mask, complement, eps_mult=1.0):
            count('xupdate')
            m = (1 - mask) if complement else mask
            return xs + m * eps * eps_mult * vs

        @staticmethod
        def su3_heads_vupdate_pair(z, pack, v, f, eps1, sign1, eps2, sign2, negate_between=False):
            count('vupdate_pair')
            v1 = v + sign1 * eps1 * (f + z.to(v.dtype))
            if negate_between:
                v1 = -v1
            v2 = v1 + sign2 * eps2 * (f + z.to(v.dtype))
            return v2, (z.double() * sign1 * eps1).sum(1) + (z.double() * sign2 * eps2).sum(1)

        @staticmethod
        def su3_update_gauge_planar_pair(xs, vs, eps, mask, first_complement, eps_mult=1.0):
            count('xupdate_pair')
            for comp in (first_complement, not first_complement):
                m = (1 - mask) if comp else mask
                xs = xs + m * eps * eps_mult * vs
            return xs

    class FakeVnet:
        def parameters(self):
            return iter([torch.zeros(1, dtype=torch.float64)])

        def hidden(self, inputs):
            return inputs[0] + 2.0 * inputs[1]

        def heads_pack(self, perm):
            return None

    monkeypatch.setattr(dmod, 'ops', FakeOps)
    masks = [(torch.arange(n) % 2 == k % 2).double() for k in range(nlf)]
    shared = FakeVnet()            # one vnet for all layers (use_separate_networks = false, conf/dynamics/su3.yaml)
    veps = torch.linspace(0.04, 0.06, nlf, dtype=torch.float64)
    xeps = torch.linspace(0.06, 0.08, nlf, dtype=torch.float64)
    fake = types.SimpleNamespace(
        config=types.SimpleNamespace(nleapfrog=nlf), _fcache=None,
        _eps_tensors=lambda: (xeps, veps), _fused_input=lambda vnet, nb_: False,
        _planar_consts=lambda: (None, masks), unflatten=lambda t: t, _get_vnet=lambda step: shared,
        compute_accept_prob=lambda s0, s1, sld: torch.exp(-sld.abs()))
    fake._reuse_force = lambda: dmod.Dynamics._reuse_force(fake)
    torch.manual_seed(0)
    st = State(torch.randn(nb, n, dtype=torch.float64), torch.randn(nb, n, dtype=torch.float64), torch.tensor(2.0))
    res = {}
    for mode, pair in (('never', 'auto'), ('always', 'never'), ('always', 'auto')):
        fake.reuse_force, fake.pair_updates = mode, pair
        calls.clear()
        out, met = dmod.Dynamics._transition_kernel_fb_planar(fake, st)
        res[(mode, pair)] = (out.x, out.v, met['acc'], met['sumlogdet'], dict(calls))
    # the reference's schedule: everything recomputed for every update, one kernel per update
    assert res[('never', 'auto')][4] == {'force': 4 * nlf, 'project': 8 * nlf, 'vupdate': 4 * nlf, 'xupdate': 4 * nlf}
    # reuse: once per distinct link configuration
    assert res[('always', 'never')][4] == {'force': 2 * nlf + 1, 'project': 2 * (2 * nlf + 1), 'vupdate': 4 * nlf,
                                           'xupdate': 4 * nlf}
    # paired passes: the first and the last momentum update stand alone, all others (turn-around included) pair up
    assert res[('always', 'auto')][4] == {'force': 2 * nlf + 1, 'project': 2 * (2 * nlf + 1), 'vupdate': 2,
                                          'vupdate_pair': 2 * nlf - 1, 'xupdate_pair': 2 * nlf}
    for key in (('always', 'never'), ('always', 'auto')):
        for a, b in zip(res[('never', 'auto')][:2], res[key][:2]):
            assert torch.equal(a, b), key              # links and momenta: the same operations in the same order
        for a, b in zip(res[('never', 'auto')][2:4], res[key][2:4]):
            assert torch.allclose(a, b, rtol=1e-14, atol=1e-14), key    # log-Jacobians are added pairwise


def test_call_vnet_packs_projected_inputs_in_the_reference_order():
    """`_call_vnet` for SU(3) (dynamics.py:1142-1159): vnet((group_to_vec(x), group_to_vec(force)))"""
    fake, calls = make('never')
    seen = {}

    class Vnet:
        def parameters(self):
            return iter([torch.zeros(1, dtype=torch.float64)])

        def __call__(self, inputs):
            seen['inputs'] = inputs
            return inputs[0], inputs[1], inputs[0] + inputs[1]
    fake._su3, fake._networks_built = True, True
    fake._get_vnet = lambda step: Vnet()
    fake.group_to_vec = lambda x, dt: (x * 3.0).to(dt)
    x, f = torch.randn(2, 4, dtype=torch.float64), torch.randn(2, 4, dtype=torch.float64)
    s, t, q = Dynamics._call_vnet(fake, 0, (x, f))
    assert torch.equal(seen['inputs'][0], x * 3.0) and torch.equal(seen['inputs'][1], f * 3.0)
    assert torch.equal(q, x * 3.0 + f * 3.0) and torch.equal(s, x * 3.0) and torch.equal(t, f * 3.0)
