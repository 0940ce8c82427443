"""CPU tier: host logic of the force / vnet-input reuse between consecutive v-updates
(`Dynamics._force`, `_vnet_vecs`; opt-in, `reuse_force = 'always'`).  The methods only need
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
