"""Kernel C: the synthetic 512 x 512 two-dimensional extension (BASELINE config 5).  The reference has
no 2-D operator, so parity is against oracle/qc_oracle2d.py (2-D analogue of phys_base, fft2-free: T_x + T_y), and
the link to the reference is the separable case: a converged 2-D step of a product state equals the outer product of
two 1-D steps of the pinned 1-D oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import qc_oracle as qo
from oracle import qc_oracle2d as q2

N = 512


@pytest.fixture(scope="module")
def ctx():
    from qcontrol_b200 import engine
    c = engine.Context(0)
    yield c
    c.close()


def _setup(L=10.0, m=0.5):
    dx = L / (N - 1)
    x = (np.arange(N) - (N - 1) / 2) * dx
    ak = qo.initak(N, dx, 2, 1) * (-qo.HART_TO_CM / (2.0 * m * qo.DALT_TO_AU))
    return dx, x, np.ascontiguousarray(ak.real)


def test_one_step_vs_oracle_random(ctx):
    """Seeded random states, a non-separable potential (coupled anharmonic wells), per-wavefunction potentials."""
    from qcontrol_b200 import engine, math_base
    rng = np.random.default_rng(77)
    dx, x, ak = _setup()
    X, Y = np.meshgrid(x, x, indexing="ij")
    batch, nch = 3, 64
    v = np.stack([3000.0 * (X ** 2 + (1.0 + 0.1 * b) * Y ** 2) + 800.0 * X * Y + 50.0 * X ** 2 * Y ** 2 for b in range(batch)])
    env = np.exp(-(X ** 2 + Y ** 2) / 2.0)
    psi = (rng.standard_normal((batch, N, N)) + 1j * rng.standard_normal((batch, N, N))) * env
    psi /= np.sqrt(dx * dx * np.sum(np.abs(psi) ** 2, axis=(1, 2), keepdims=True))
    emax = float(v.max() + 2 * abs(ak[N // 2 - 1]))
    emin = 0.0
    dt = 1.0e-18
    t_sc = dt * (emax - emin) * qo.CM_TO_ERG / 4.0 / qo.RED_PLANCK_H
    pl = engine.Plan2D(ctx, N, nch, batch)
    pl.set_static(v, ak, ak, dx * dx)
    xp, dv = math_base.newton_table(nch, t_sc)
    pl.set_poly(xp, dv, emin, emax, t_sc)
    pl.set_state(psi)
    pl.sweep(1, renorm=True)
    out = pl.get_state()
    norms = pl.get_norms()
    for b in range(batch):
        ref, nrm = q2.step_2d(psi[b], v[b], ak.astype(complex), ak.astype(complex), nch, dt, dx * dx, emin, emax)
        err = float(np.sqrt(dx * dx * np.sum(np.abs(out[b] - ref) ** 2)))
        assert err <= 1e-10, (b, err)
        assert abs(norms[b] - nrm) <= 1e-10
    # several steps in one launch == the same steps one launch at a time, bit for bit
    pl.set_state(psi)
    pl.sweep(3, renorm=True)
    a = pl.get_state()
    pl.set_state(psi)
    for _ in range(3):
        pl.sweep(1, renorm=True)
    assert np.array_equal(a, pl.get_state())
    pl.close()


def test_separable_product_state_matches_the_1d_path(ctx):
    """V(x,y) = V1(x) + V2(y), psi = f(x) g(y): the 2-D kernel against the outer product of two 1-D propagations by
    the pinned 1-D oracle (= the reference's phys_base.prop_cpu), 20 converged steps."""
    from qcontrol_b200 import engine, math_base
    dx, x, ak = _setup()
    akc = ak.astype(complex)
    k1, k2 = 9000.0, 5000.0
    v1, v2 = k1 * x ** 2 / 2, k2 * (x - 0.3) ** 2 / 2
    f = np.exp(-(x - 0.5) ** 2 / 0.8).astype(complex)
    g = (np.exp(-(x + 0.4) ** 2 / 1.1) * np.exp(1j * 2.0 * x)).astype(complex)
    f /= np.sqrt(dx * np.sum(np.abs(f) ** 2))
    g /= np.sqrt(dx * np.sum(np.abs(g) ** 2))
    nch, dt, steps = 64, 4.0e-18, 20
    v = v1[:, None] + v2[None, :]
    emax2, emin2 = q2.energy_window_2d(v, akc, akc, N)
    t_sc2 = dt * (emax2 - emin2) * qo.CM_TO_ERG / 4.0 / qo.RED_PLANCK_H
    pl = engine.Plan2D(ctx, N, nch, 1)
    pl.set_static(v, ak, ak, dx * dx)
    xp, dv = math_base.newton_table(nch, t_sc2)
    pl.set_poly(xp, dv, emin2, emax2, t_sc2)
    pl.set_state((f[:, None] * g[None, :])[None])
    pl.sweep(steps, renorm=True)
    out = pl.get_state()[0]

    def prop_1d(psi, vv):
        pb = qo.Problem(hamil="ntriv", ntriv=1, np_=N, nlevs=2, nb=1, dx=dx, v=np.stack([vv, vv]), vmin=np.zeros(2),
                        akx2=akc, nch=nch)
        emax, emin = float(vv.max() + abs(ak[N // 2 - 1])), 0.0
        t_sc = dt * (emax - emin) * qo.CM_TO_ERG / 4.0 / qo.RED_PLANCK_H
        st = np.stack([psi, np.zeros_like(psi)])
        for _ in range(steps):
            st = qo.newton_propagate(pb, st, t_sc, emin, emax, np.float64(0.0), 0.0, np.float64(0.0))
            st = st / np.sqrt(abs(np.vdot(st, st) * dx))
        return st[0]

    ref = prop_1d(f, v1)[:, None] * prop_1d(g, v2)[None, :]
    err = float(np.sqrt(dx * dx * np.sum(np.abs(out - ref) ** 2)))
    assert err <= 1e-9, err
    pl.close()


def test_abi_errors_2d(ctx):
    from qcontrol_b200 import engine
    from qcontrol_b200._lib import QcError
    with pytest.raises(QcError) as e:
        engine.Plan2D(ctx, 256, 64, 1)
    assert "512" in str(e.value)
    pl = engine.Plan2D(ctx, N, 64, 1)
    with pytest.raises(QcError) as e:
        pl.sweep(1)
    assert e.value.code == -3
    pl.close()
