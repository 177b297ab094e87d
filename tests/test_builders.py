"""Filter-instance construction: the product library's host builders against the oracle's independent ones and
against textbook values."""
import numpy as np
import pytest


def test_sg_textbook_values(L, O):
    for B in (L.LibBuilders(), O.OracleBuilders()):
        assert np.allclose(B.sg_coeffs(5, 3, 1), np.array([1, -8, 0, 8, -1]) / 12.0, atol=1e-14)
        assert np.allclose(B.sg_coeffs(7, 3, 1), np.array([22, -67, -58, 0, 58, 67, -22]) / 252.0, atol=1e-14)
        assert np.allclose(B.sg_coeffs(5, 2, 0), np.array([-3, 12, 17, 12, -3]) / 35.0, atol=1e-14)


@pytest.mark.parametrize("n,deg", [(6, 3), (44, 3), (5, 1), (64, 3), (9, 2)])
def test_lsq_fit_matrix_reproduces_polynomials(L, O, n, deg):
    x = np.arange(n, dtype=float)
    rng = np.random.default_rng(n)
    c = rng.normal(size=deg + 1)
    y = sum(c[j] * x ** j for j in range(deg + 1))
    Al, Ao = L.LibBuilders().lsq_fit_matrix(n, deg), O.OracleBuilders().lsq_fit_matrix(n, deg)
    assert np.allclose(y @ Al, c, rtol=1e-8, atol=1e-8)
    assert np.allclose(Al, Ao, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("sigma,flat,L_,zac", [(312.5, 156, 2375, 0), (312.5, 156, 2375, 1), (62.5, 62, 2375, 1),
                                               (80.5, 40, 601, 1), (1000.0, 250, 2375, 0)])
def test_cuspzac_coeffs_agree_and_have_the_expected_shape(L, O, sigma, flat, L_, zac):
    tau = 6.25e8
    fl = (L.LibBuilders().zac_coeffs if zac else L.LibBuilders().cusp_coeffs)(sigma, flat, tau, L_, float(L_))
    fo = (O.OracleBuilders().zac_coeffs if zac else O.OracleBuilders().cusp_coeffs)(sigma, flat, tau, L_, float(L_))
    assert np.allclose(fl, fo, rtol=1e-10, atol=1e-15)
    shape = np.cumsum(fl)            # tau -> inf: coeffs = first difference of the shape
    lt = (L_ - flat) // 2
    assert np.allclose(shape[lt:lt + flat + 1], 1.0, atol=1e-5)   # unit flat top for beta = L
    if zac:
        assert abs(shape.sum()) < 1e-6 * np.abs(shape).sum()      # zero area


def test_builder_argument_errors(L):
    lib = L.load_library()
    import ctypes as C
    buf = (C.c_double * 16)()
    assert lib.lgdsp_sg_coeffs(4, 3, 1, buf) != 0      # even length
    assert lib.lgdsp_sg_coeffs(3, 1, 2, buf) != 0      # derivative above the degree
    assert lib.lgdsp_lsq_fit_matrix(3, 3, buf) != 0
    assert lib.lgdsp_cusp_coeffs(-1.0, 2, 1.0, 16, 1.0, buf) != 0


def test_sg_underdetermined_minimum_norm(L, O):
    """3 taps / degree 3 (the in-trace filter of dsp_icpc_compressed with the example config, src/dsp_icpc.jl:439):
    minimum-norm solution = pseudo-inverse row, in the library and in the oracle"""
    x = np.arange(-1, 2, dtype=np.float64)
    V = np.vander(x, 4, increasing=True)
    ref = np.linalg.pinv(V)[1]
    assert np.allclose(L.LibBuilders().sg_coeffs(3, 3, 1), ref, atol=1e-14)
    assert np.allclose(O.OracleBuilders().sg_coeffs(3, 3, 1), ref, atol=1e-14)
    assert np.allclose(ref, [-0.25, 0.0, 0.25])
    x = np.arange(-2, 3, dtype=np.float64)
    ref5 = np.linalg.pinv(np.vander(x, 7, increasing=True))[1]
    assert np.allclose(L.LibBuilders().sg_coeffs(5, 6, 1), ref5, atol=1e-12)
    assert np.allclose(O.OracleBuilders().sg_coeffs(5, 6, 1), ref5, atol=1e-12)


def test_against_independent_library_implementations(L, O):
    """the published algorithms behind the RadiationDetectorDSP primitives, from independent implementations (scipy / numpy):
    Savitzky-Golay derivative coefficients, the inverse-CR biquad, the integrator, valid-mode FIR and the trapezoid"""
    import scipy.signal as sig
    rng = np.random.default_rng(9)
    for n_taps, deg in ((5, 3), (7, 3), (13, 3), (7, 2), (21, 4)):
        ref = sig.savgol_coeffs(n_taps, deg, deriv=1, use="dot")          # s[j] = sum_k h[k] y[j+k], per-sample derivative
        assert np.allclose(L.LibBuilders().sg_coeffs(n_taps, deg, 1), ref, atol=1e-13), (n_taps, deg)
        assert np.allclose(O.OracleBuilders().sg_coeffs(n_taps, deg, 1), ref, atol=1e-12), (n_taps, deg)
    y = rng.normal(0, 1, 500)
    # InvCRFilter(tau): biquad b = (1/alpha, -1), a = (1, -1), alpha = RC/(RC+1)   
    RC = 31250.0
    alpha = RC / (RC + 1.0)
    assert np.allclose(O.invcr(y, 1.0 / alpha - 1.0), sig.lfilter([1.0 / alpha, -1.0], [1.0, -1.0], y), rtol=1e-12, atol=1e-12)
    assert np.allclose(O.integrator(y), np.cumsum(y), rtol=1e-12, atol=1e-12)
    c = rng.normal(0, 1, 37)
    assert np.allclose(O.fir_valid(y, c), np.convolve(y, c, mode="valid"), rtol=1e-12, atol=1e-12)
    h = np.array([0.2, -0.5, 0.1, 0.7, -0.3])
    assert np.allclose(O.corr_valid(y, h), np.correlate(y, h, mode="valid"), rtol=1e-12, atol=1e-12)
    # TrapezoidalChargeFilter(avg, gap, avg2) as an FIR: +1/avg2 over the trailing window, -1/avg over the leading one
    a, g, a2 = 6, 3, 8
    k = np.concatenate([np.full(a2, 1.0 / a2), np.zeros(g), np.full(a, -1.0 / a)])
    assert np.allclose(O.trap(y, a, g, a2), np.convolve(y, k, mode="valid"), rtol=1e-12, atol=1e-12)
