"""Device-side collocation / test point generation (scasml_geometry_points): the NumPy restatement's properties on the CPU,
bit-exact agreement of the CUDA kernel with it on the GPU, and the fit / evaluation API taking device-generated points as they are."""
import numpy as np
import pytest

from oracle.equation import EquationOracle


@pytest.mark.parametrize("d", [6, 20, 100])
def test_philox_geometry_restatement_properties(d):
    eq = EquationOracle(d + 1)
    dom, bdy = eq.generate_data_philox(300, 90, seed=7)
    assert dom.shape == (300, d + 1) and bdy.shape == (90, d + 1)
    for pts in (dom, bdy):
        assert np.array_equal(pts, pts.astype(np.float16).astype(np.float64))            # float16-valued (DeepXDE float16)
        assert np.all(np.abs(pts[:, :d]) <= 0.5) and np.all((pts[:, d] >= 0.0) & (pts[:, d] <= 0.5))
    assert np.all(np.sum(np.abs(bdy[:, :d]) == 0.5, axis=1) >= 1)                       # every boundary point lies on a face
    assert np.mean(np.sum(np.abs(dom[:, :d]) == 0.5, axis=1)) < 0.1 * d                  # domain points (almost) never do
    again = eq.generate_data_philox(300, 90, seed=7)
    assert np.array_equal(dom, again[0]) and np.array_equal(bdy, again[1])
    other = eq.generate_data_philox(300, 90, seed=8)
    assert not np.array_equal(dom, other[0])
    # prefix property of a counter-based stream: fewer points are a prefix of more points
    assert np.array_equal(eq.generate_data_philox(100, 30, seed=7)[0], dom[:100])
    # first moments of the uniform law
    assert abs(dom[:, :d].mean()) < 0.02 and abs(dom[:, d].mean() - 0.25) < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize("d", [6, 33, 100])
def test_device_geometry_matches_restatement_bit_for_bit(d):
    from tests.test_gpu_parity import _product
    P = _product()
    eq_o = EquationOracle(d + 1)
    eq = P["Eq"](d + 1)
    for seed in (1234, 5):
        dom, bdy = eq.generate_data_device(257, 65, seed=seed)
        want_dom, want_bdy = eq_o.generate_data_philox(257, 65, seed=seed)
        assert np.array_equal(dom.cpu().numpy(), want_dom)
        assert np.array_equal(bdy.cpu().numpy(), want_bdy)
    tdom, tbdy = eq.generate_test_data_device(40, 8, seed=42)
    want = eq_o.generate_data_philox(40, 8, seed=42)
    assert np.array_equal(tdom.cpu().numpy(), want[0]) and np.array_equal(tbdy.cpu().numpy(), want[1])


@pytest.mark.gpu
def test_fit_and_solve_take_device_generated_points():
    from tests.test_gpu_parity import _product
    P = _product()
    d = 6
    eq = P["Eq"](d + 1)
    dom, bdy = eq.generate_data_device(40, 12, seed=3)
    idx = np.random.default_rng(7).choice(d, 5, replace=False)
    sol0 = np.random.default_rng(0).standard_normal(3 * 40) * 1e-3
    gp_dev = P["GP"](eq, idx_set=idx)
    gp_dev.GPsolver(dom, bdy, sol0=sol0)                                  # CUDA tensors in
    gp_host = P["GP"](eq, idx_set=idx)
    gp_host.GPsolver(dom.cpu().numpy(), bdy.cpu().numpy(), sol0=sol0)     # the same points as host arrays
    assert np.array_equal(gp_dev.right_vector, gp_host.right_vector)
    X = eq.generate_test_data_device(10, 2, seed=42)[0]
    assert np.array_equal(gp_dev.predict(X), gp_host.predict(X.cpu().numpy()))
    s = P["ScaSML"](eq, gp_dev)
    a = s.u_solve(2, 2, X)
    s2 = P["ScaSML"](eq, gp_host)
    b = s2.u_solve(2, 2, X.cpu().numpy())
    assert np.array_equal(a, b)
