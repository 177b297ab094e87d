"""Host-side configuration surface of the hot path.

Mirrors `DSPConfig` (/root/reference/src/types.jl:32-93), `_create_dsp_config` (src/utils.jl:14-70) and
`get_fltpars` (src/utils.jl:72-82), and resolves every time quantity into the sample-domain
`lgdsp_icpc_params` of include/lgdsp_b200.h with the reference's own expressions:

    from  = round(Int, ustrip(NoUnits, (start - first_x) / step_x)) + firstindex      src/tailstats.jl:16-18
    navg  = round(Int, ustrip(NoUnits, avgtime / step))                               [RDDSP fltinstance]
    min_n = max(1, round(Int, ustrip(NoUnits, mintot / step)))                        src/intersect_maximum.jl:20

Julia's round() is ties-to-even and Unitful promotes mixed units to seconds before subtracting, which decides
the exact ties of the reference's example config; `Q`, `_sub_over_step` and `_ratio`
below emulate that evaluation order.  (The Julia wrapper julia/LegendDSPB200.jl evaluates the genuine
expressions instead.)

Policies for RadiationDetectorDSP semantics that the reference's tests do not pin ("parity unpinned",
DESIGN.md section 3) are explicit switches of `RddspPolicy`.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import math
from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

from . import _abi

# ----------------------------------------------------------------------------------------------
# minimal quantity emulation (only what the reference's index arithmetic needs)
# ----------------------------------------------------------------------------------------------
# exact rational factors to seconds, as Unitful holds them
_TO_S = {"ns": (1, 10**9), "us": (1, 10**6), "µs": (1, 10**6), "μs": (1, 10**6), "ms": (1, 10**3), "s": (1, 1)}


@dataclass(frozen=True)
class Q:
    """value * unit, unit in {ns, us, ms, s}"""
    val: float
    unit: str = "ns"

    def __post_init__(self):
        if self.unit not in _TO_S:
            raise ValueError(f"unknown time unit {self.unit!r}")

    def to_seconds(self) -> float:
        num, den = _TO_S[self.unit]
        # Float64 * Rational promotes the rational to Float64 first (Base rational.jl)
        return self.val * (num / den)

    def ns(self) -> float:
        """plain conversion to nanoseconds (for quantities that are not index-rounded)"""
        num, den = _TO_S[self.unit]
        f = num * 10**9 // den if (num * 10**9) % den == 0 else num * 1e9 / den
        return self.val * f

    def __add__(self, o: "Q") -> "Q":
        if self.unit == o.unit:
            return Q(self.val + o.val, self.unit)
        return Q(self.to_seconds() + o.to_seconds(), "s")

    def __sub__(self, o: "Q") -> "Q":
        if self.unit == o.unit:
            return Q(self.val - o.val, self.unit)
        return Q(self.to_seconds() - o.to_seconds(), "s")

    def __mul__(self, f: float) -> "Q":
        return Q(self.val * f, self.unit)

    __rmul__ = __mul__

    def __truediv__(self, f: float) -> "Q":
        return Q(self.val / f, self.unit)


def us(v: float) -> Q:
    return Q(float(v), "us")


def ns(v: float) -> Q:
    return Q(float(v), "ns")


def _ratio(a: Q, b: Q) -> float:
    """ustrip(NoUnits, a / b): divide the values, then apply the exact unit conversion factor"""
    r = a.val / b.val
    na, da = _TO_S[a.unit]
    nb, db = _TO_S[b.unit]
    num, den = na * db, da * nb
    g = math.gcd(num, den)
    num //= g
    den //= g
    if den == 1:
        return r * num
    return r * (num / den)


def julia_round(x: float) -> int:
    """round(Int, x): IEEE ties-to-even"""
    return int(np.rint(x))


def _sub_over_step(t: Q, first_x: Q, step_x: Q) -> int:
    """round(Int, ustrip(NoUnits, (t - first_x)/step_x))  -- 0-based sample index of time t"""
    return julia_round(_ratio(t - first_x, step_x))


# ----------------------------------------------------------------------------------------------
# policies for the un-pinned RadiationDetectorDSP semantics
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class RddspPolicy:
    # SavitzkyGolayFilter: kernel length when round(length/dt) is even: "up" -> n+1, "down" -> n-1
    sg_even_length: str = "up"
    # time axis of the valid-mode SG trace: "center" -> trace j <-> sample j+(n_taps-1)/2; "trailing" -> j+n_taps-1
    sg_time_axis: str = "center"
    # CUSP/ZAC normalisation: "beta_over_len" -> coeffs*beta/L (unit flat-top gain for beta = L); "beta" -> coeffs*beta
    cuspzac_norm: str = "beta_over_len"


DEFAULT_POLICY = RddspPolicy()


# ----------------------------------------------------------------------------------------------
# DSPConfig   (src/types.jl:32-93)
# ----------------------------------------------------------------------------------------------
@dataclass
class DSPConfig:
    enc_pickoff_trap: Q
    enc_pickoff_zac: Q
    enc_pickoff_cusp: Q
    flt_length_cusp: Q
    flt_length_zac: Q
    t0_threshold: float
    inTraceCut_std_threshold: float
    sg_flt_degree: int
    bl_window: Tuple[Q, Q]
    tail_window: Tuple[Q, Q]
    current_window: Tuple[Q, Q]
    qdrift_int_length: Tuple[Q, Q, Q]       # (first, step, last)
    lq_int_length: Tuple[Q, Q, Q]
    e_grid_rt_trap: Tuple[Q, Q, Q]          # (start, step, stop)
    e_grid_ft_trap: Tuple[Q, Q, Q]
    e_grid_rt_zac: Tuple[Q, Q, Q]
    e_grid_ft_zac: Tuple[Q, Q, Q]
    e_grid_rt_cusp: Tuple[Q, Q, Q]
    e_grid_ft_cusp: Tuple[Q, Q, Q]
    a_grid_wl_sg: Tuple[Q, Q, Q]
    default_flt_param: Dict[str, Any]
    kwargs_pars: Dict[str, Any]
    auxbl1_window: Tuple[Q, Q] = (us(0.0), us(20.0))
    auxbl2_window: Tuple[Q, Q] = (us(20.0), us(39.0))
    auxpz1_window: Tuple[Q, Q] = (us(70.0), us(90.0))
    auxpz2_window: Tuple[Q, Q] = (us(90.0), us(110.0))

    @staticmethod
    def from_dict(pd: Dict[str, Any]) -> "DSPConfig":
        """_create_dsp_config(dsp_metadata)  src/utils.jl:14-70; `pd` uses the same keys as the PropDict"""
        def win(k):
            return (pd[k]["min"], pd[k]["max"])

        def grid(d):
            return (d["start"], d["step"], d["stop"])

        q = pd["qdrift_int_length"]
        l = pd["lq_int_length"]
        # src/utils.jl:40-42: the step is 0.1*unit(first(qdrift_int_length)) for BOTH ranges (sic)
        qstep = Q(0.1, q[0].unit)
        return DSPConfig(
            enc_pickoff_trap=pd["enc_pickoff_trap"], enc_pickoff_zac=pd["enc_pickoff_zac"],
            enc_pickoff_cusp=pd["enc_pickoff_cusp"],
            flt_length_cusp=pd["flt_length_cusp"], flt_length_zac=pd["flt_length_zac"],
            t0_threshold=float(pd["t0_threshold"]),
            inTraceCut_std_threshold=float(pd["inTraceCut_std_threshold"]),
            sg_flt_degree=int(pd["sg_flt_degree"]),
            bl_window=win("bl_window"), tail_window=win("tail_window"), current_window=win("current_window"),
            qdrift_int_length=(q[0], qstep, q[-1]), lq_int_length=(l[0], qstep, l[-1]),
            e_grid_rt_trap=grid(pd["e_grid_trap"]["rt"]), e_grid_ft_trap=grid(pd["e_grid_trap"]["ft"]),
            e_grid_rt_zac=grid(pd["e_grid_zac"]["rt"]), e_grid_ft_zac=grid(pd["e_grid_zac"]["ft"]),
            e_grid_rt_cusp=grid(pd["e_grid_cusp"]["rt"]), e_grid_ft_cusp=grid(pd["e_grid_cusp"]["ft"]),
            a_grid_wl_sg=grid(pd["a_grid_wl_sg"]),
            default_flt_param=pd["flt_defaults"], kwargs_pars=pd["kwargs_pars"],
            auxbl1_window=win("auxbl1_window"), auxbl2_window=win("auxbl2_window"),
            auxpz1_window=win("auxpz1_window"), auxpz2_window=win("auxpz2_window"),
        )


def grid_values(g: Tuple[Q, Q, Q]) -> list:
    """collect(start:step:stop) -- Julia range semantics (length = floor((stop-start)/step + eps) + 1)"""
    start, step, stop = g
    n = int(math.floor((stop.val - start.val) / step.val + 1e-9)) + 1
    return [Q(start.val + i * step.val, start.unit) for i in range(n)]


def example_config_dict() -> Dict[str, Any]:
    """make_fake_config()  /root/reference/test/test_dsp_icpc.jl:50-161, as data"""
    rtft = {"rt": {"start": us(1.0), "stop": us(16.0), "step": us(0.5)},
            "ft": {"start": us(1.0), "stop": us(4.0), "step": us(0.2)}}
    return {
        "enc_pickoff_trap": us(40.0), "enc_pickoff_zac": us(41.0), "enc_pickoff_cusp": us(41.0),
        "bl_window": {"min": us(0.0), "max": us(39.0)},
        "tail_window": {"min": us(70.0), "max": us(110.0)},
        "current_window": {"min": us(43.0), "max": us(62.0)},
        "auxbl1_window": {"min": us(0.0), "max": us(20.0)},
        "auxbl2_window": {"min": us(20.0), "max": us(39.0)},
        "auxpz1_window": {"min": us(70.0), "max": us(90.0)},
        "auxpz2_window": {"min": us(90.0), "max": us(110.0)},
        "flt_length_cusp": us(38.0), "flt_length_zac": us(38.0),
        "t0_threshold": 4.0, "inTraceCut_std_threshold": 5, "sg_flt_degree": 3,
        "qdrift_int_length": (us(2.5), us(5.0)), "lq_int_length": (us(2.5), us(5.0)),
        "e_grid_trap": rtft, "e_grid_zac": rtft, "e_grid_cusp": rtft,
        "a_grid_wl_sg": {"start": ns(30.0), "stop": ns(350.0), "step": ns(32.0)},
        "flt_defaults": {
            "sg": ns(100.0),
            "trap": {"rt": us(5.0), "ft": us(2.5)},
            "zac": {"rt": us(5.0), "ft": us(2.5)},
            "cusp": {"rt": us(5.0), "ft": us(2.5)},
        },
        "kwargs_pars": {
            "fc_bit_depth": 16,
            "t0_flt_pars": [ns(40.0), ns(100.0), ns(2000.0)],
            "t0_mintot": ns(1500.0), "tx_mintot": ns(32.0), "intrace_mintot": ns(100.0),
            "int_interpolation_order": 3, "int_interpolation_length": ns(100.0),
            "sig_interpolation_order": 3, "sig_interpolation_length": ns(700.0),
        },
    }


def example_config() -> DSPConfig:
    return DSPConfig.from_dict(example_config_dict())


def tiefree_config() -> DSPConfig:
    """The example config with its window edges moved onto multiples of 16 ns (no rounding ties):
    used for benchmarks and bit-exact index checks."""
    d = example_config_dict()
    d["bl_window"] = {"min": us(0.0), "max": us(39.008)}        # 2438 samples -> idx 0..2438
    d["current_window"] = {"min": us(43.008), "max": us(62.0)}  # 2688 .. 3875
    return DSPConfig.from_dict(d)


def get_fltpars(pars_filter: Optional[Dict[str, Any]], flt: str, cfg: DSPConfig):
    """get_fltpars(pd, flt, dspconfig)  src/utils.jl:72-82"""
    pd = pars_filter or {}
    if flt == "sg":
        return pd.get("sg", {}).get("wl", cfg.default_flt_param["sg"])
    if flt not in pd:
        return cfg.default_flt_param[flt]["rt"], cfg.default_flt_param[flt]["ft"]
    return (pd[flt].get("rt", cfg.default_flt_param[flt]["rt"]),
            pd[flt].get("ft", cfg.default_flt_param[flt]["ft"]))


# ----------------------------------------------------------------------------------------------
# builders: where the filter coefficients come from (product library by default; the oracle's
# independent implementations in tests)
# ----------------------------------------------------------------------------------------------
class LibBuilders:
    """coefficient construction through the product library's host-side C functions"""

    def __init__(self, lib=None):
        if lib is None:
            from ._lib import load_library
            lib = load_library()
        self._lib = lib

    def lsq_fit_matrix(self, n: int, degree: int) -> np.ndarray:
        A = np.zeros((n, degree + 1), dtype=np.float64)
        rc = self._lib.lgdsp_lsq_fit_matrix(n, degree, A.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise ValueError(f"lgdsp_lsq_fit_matrix({n}, {degree}) failed: {rc}")
        return A

    def sg_coeffs(self, n_taps: int, degree: int, derivative: int) -> np.ndarray:
        h = np.zeros(n_taps, dtype=np.float64)
        rc = self._lib.lgdsp_sg_coeffs(n_taps, degree, derivative, h.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise ValueError(f"lgdsp_sg_coeffs({n_taps}, {degree}, {derivative}) failed: {rc}")
        return h

    def cusp_coeffs(self, sigma, flat, tau, L, beta) -> np.ndarray:
        c = np.zeros(L, dtype=np.float64)
        rc = self._lib.lgdsp_cusp_coeffs(sigma, flat, tau, L, beta, c.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise ValueError(f"lgdsp_cusp_coeffs failed: {rc}")
        return c

    def zac_coeffs(self, sigma, flat, tau, L, beta) -> np.ndarray:
        c = np.zeros(L, dtype=np.float64)
        rc = self._lib.lgdsp_zac_coeffs(sigma, flat, tau, L, beta, c.ctypes.data_as(C.POINTER(C.c_double)))
        if rc != 0:
            raise ValueError(f"lgdsp_zac_coeffs failed: {rc}")
        return c


# ----------------------------------------------------------------------------------------------
# resolution into sample units
# ----------------------------------------------------------------------------------------------
def _trap(avg: Q, gap: Q, step: Q, avg2: Optional[Q] = None) -> _abi.Trap:
    navg = julia_round(_ratio(avg, step))
    ngap = julia_round(_ratio(gap, step))
    navg2 = navg if avg2 is None else julia_round(_ratio(avg2, step))
    return _abi.Trap(navg, ngap, navg2, 0)


def _min_n(mintot: Q, step: Q) -> int:
    return max(1, julia_round(_ratio(mintot, step)))


def _fill_dni(d: _abi.Dni, degree: int, length: Q, step: Q, builders) -> None:
    n_w = julia_round(_ratio(length, step))
    if not (degree + 1 <= n_w <= _abi.LGDSP_MAX_DNI) or degree > _abi.LGDSP_MAX_DNI_DEG:
        raise ValueError(f"PolynomialDNI(degree={degree}, n_w={n_w}) outside the supported range")
    A = builders.lsq_fit_matrix(n_w, degree)
    d.n_w, d.degree = n_w, degree
    flat = A.reshape(-1)
    for i, v in enumerate(flat):
        d.A[i] = float(v)


def _sg_taps(length: Q, step: Q, policy: RddspPolicy) -> int:
    n = julia_round(_ratio(length, step))
    if n % 2 == 0:
        n = n + 1 if policy.sg_even_length == "up" else n - 1
    return n


def _fill_sg(s: _abi.Sg, length: Q, degree: int, step: Q, policy: RddspPolicy, builders) -> None:
    n = _sg_taps(length, step, policy)
    # n <= degree (underdetermined fit) gives the minimum-norm coefficients, see lgdsp_sg_coeffs
    if n > _abi.LGDSP_MAX_SG or n < 1:
        raise ValueError(f"SavitzkyGolayFilter with {n} taps / degree {degree} outside the supported range")
    h = builders.sg_coeffs(n, degree, 1)
    s.n_taps = n
    s.offset = (n - 1) // 2 if policy.sg_time_axis == "center" else n - 1
    for i, v in enumerate(h):
        s.h[i] = float(v)


def _fill_cuspzac(cz: _abi.CuspZac, kind: str, rt: Q, ft: Q, tau: Q, length: Q, scale: float, step: Q,
                  policy: RddspPolicy, builders) -> None:
    L = julia_round(_ratio(length, step))
    flat = julia_round(_ratio(ft, step))
    sigma = _ratio(rt, step)
    tau_s = _ratio(tau, step)
    if L > _abi.LGDSP_MAX_FIR:
        raise ValueError(f"{kind} filter length {L} exceeds LGDSP_MAX_FIR")
    # the builders implement the "beta_over_len" normalisation; "beta" rescales beta by L
    beta = scale if policy.cuspzac_norm == "beta_over_len" else scale * L
    c = builders.cusp_coeffs(sigma, flat, tau_s, L, beta) if kind == "cusp" else \
        builders.zac_coeffs(sigma, flat, tau_s, L, beta)
    cz.n_taps, cz.flat, cz.sigma, cz.tau, cz.beta = L, flat, sigma, tau_s, beta
    for i, v in enumerate(c):
        cz.coeffs[i] = float(v)


def resolve_icpc_params(cfg: DSPConfig, tau: Q, pars_filter: Optional[Dict[str, Any]] = None, *,
                        n_samples: int = 8192, t_first: Q = ns(0.0), step: Q = ns(16.0),
                        groups: int = _abi.GROUP_ALL, policy: RddspPolicy = DEFAULT_POLICY,
                        builders=None, cuspzac_direct: bool = False, role: str = "full",
                        presum_rate: int = 1) -> _abi.IcpcParams:
    """Everything `dsp_icpc` (src/dsp_icpc.jl:62-230) derives from (config, tau, pars_filter) and the time axis
    of the first waveform, expressed in samples / nanoseconds.

    `role` selects one of the two passes of `dsp_icpc_compressed` (src/dsp_icpc.jl:293-499): "pre" resolves what the
    presummed waveform needs (saturation with sat_high * presum_rate :334, windows, energy filters, the in-trace filter
    SavitzkyGolayFilter(sg_wl * presum_rate / 2) :439), "wdw" what the windowed waveform needs (t0, t10..t99, Q-drift,
    currents).  What a pass does not use is filled with the smallest valid placeholder (its groups are switched off or
    its columns are not read).  "puls" is the pulser chain (`dsp_puls`, src/dsp_puls.jl:29-66): baseline window, t50 and the
    (10 us, 4 us) trapezoid only -- everything else is a placeholder, so that presummed time axes (64 / 128 ns steps, on
    which the 60 ns / 100 ns filters of the full chain have no samples) resolve."""
    if builders is None:
        builders = LibBuilders()
    if role not in ("full", "pre", "wdw", "puls", "decay"):
        raise ValueError(f"unknown role {role!r}")
    if role == "pre":
        groups = _abi.GROUP_BASE | _abi.GROUP_TIMING | _abi.GROUP_TRAPS | _abi.GROUP_CUSPZAC | _abi.GROUP_INTRACE
    elif role == "wdw":
        groups = _abi.GROUP_BASE | _abi.GROUP_TIMING | _abi.GROUP_QDRIFT | _abi.GROUP_CURRENT
    elif role == "puls":
        groups = _abi.GROUP_BASE | _abi.GROUP_TIMING | _abi.GROUP_TRAPS
    elif role == "decay":    # dsp_decay_times (src/dsp_decaytime.jl:11-25): baseline and tail windows only
        groups = _abi.GROUP_BASE
    lite = role in ("puls", "decay")    # everything the full chain adds is a placeholder
    dummy_trap = _abi.Trap(1, 0, 1, 0)
    kw = cfg.kwargs_pars
    P = _abi.IcpcParams()
    P.struct_size = C.sizeof(_abi.IcpcParams)
    P.version = _abi.LGDSP_PARAMS_VERSION
    P.n_samples = int(n_samples)
    P.groups = int(groups)
    P.t_first_ns = t_first.ns()
    P.dt_ns = step.ns()
    n = int(n_samples)

    # src/dsp_icpc.jl:93-94  (sic: 2^bit_depth - bit_depth)
    bit_depth = int(kw["fc_bit_depth"])
    P.sat_low, P.sat_high = 0, (2 ** bit_depth - bit_depth) * (int(presum_rate) if role == "pre" else 1)   # :334

    def window(w, first_x: Q, n_trace: int, what: str):
        a = _sub_over_step(w[0], first_x, step)
        b = _sub_over_step(w[1], first_x, step)
        # @assert firstindex(X) <= first(idxs) <= last(idxs) <= lastindex(X)   src/tailstats.jl:23-25
        if not (0 <= a <= b <= n_trace - 1):
            raise AssertionError(f"{what}: index range {a + 1}:{b + 1} outside 1:{n_trace}")
        return a, b

    if role == "wdw":
        # the windowed waveform is shifted by the presummed baseline (:350); its own statistics are not read
        P.bl_from, P.bl_until = 0, min(n - 1, 15)
        P.tail_from, P.tail_until = max(0, n - 16), n - 1
    else:
        P.bl_from, P.bl_until = window(cfg.bl_window, t_first, n, "bl_window")
        if role == "puls":      # no tail statistics in dsp_puls
            P.tail_from, P.tail_until = max(0, n - 16), n - 1
        else:
            P.tail_from, P.tail_until = window(cfg.tail_window, t_first, n, "tail_window")

    # InvCRFilter(tau) [RDDSP]: RC = tau/dt, alpha = RC/(RC+1), k = 1/alpha
    RC = _ratio(tau, step)
    alpha = RC / (RC + 1.0)
    P.pz_km1 = 1.0 / alpha - 1.0

    # get_t0: src/dsp_routines.jl:9-25
    fp = kw["t0_flt_pars"]
    if role == "pre" or lite:
        P.t0_trap = P.t0inv_trap = dummy_trap     # t0 / t0_inv come from the windowed waveform (:378, :458)
    else:
        P.t0_trap = _trap(fp[0], fp[1], step, fp[2])
        P.t0inv_trap = _trap(ns(40.0), ns(100.0), step, ns(2000.0))   # default flt_pars, src/dsp_icpc.jl:207
    # the presummed pass has no t0 (its trapezoid is a placeholder): a threshold nothing reaches keeps the crossing search idle
    P.t0_threshold = 1e300 if (role == "pre" or lite) else float(cfg.t0_threshold)
    P.t0_min_n = _min_n(kw["t0_mintot"], step)
    P.tx_min_n = _min_n(kw["tx_mintot"], step)
    for i, f in enumerate((0.1, 0.5, 0.8, 0.9, 0.99)):
        P.tx_frac[i] = f

    # get_qdrift: only first(Δt) and last(Δt) are used (src/dsp_routines.jl:59-60)
    P.qdrift_first_ns = cfg.qdrift_int_length[0].ns()
    P.qdrift_last_ns = cfg.qdrift_int_length[2].ns()
    P.lq_first_ns = cfg.lq_int_length[0].ns()
    P.lq_last_ns = cfg.lq_int_length[2].ns()
    if role == "pre" or lite:   # Q-drift is evaluated on the windowed waveform only (:391-394)
        _fill_dni(P.int_dni, 1, step * 2.0, step, builders)
    else:
        _fill_dni(P.int_dni, int(kw["int_interpolation_order"]), kw["int_interpolation_length"], step, builders)
    if role == "wdw" or lite:   # the SignalEstimator of the energies runs on the presummed waveform only (:407-428)
        _fill_dni(P.sig_dni, 1, step * 2.0, step, builders)
    else:
        _fill_dni(P.sig_dni, int(kw["sig_interpolation_order"]), kw["sig_interpolation_length"], step, builders)

    # src/dsp_icpc.jl:147-160
    trap_rt, trap_ft = get_fltpars(pars_filter, "trap", cfg)
    cusp_rt, cusp_ft = get_fltpars(pars_filter, "cusp", cfg)
    zac_rt, zac_ft = get_fltpars(pars_filter, "zac", cfg)
    sg_wl = get_fltpars(pars_filter, "sg", cfg)
    if role in ("wdw", "decay"):
        P.trap_10410 = P.trap_535 = P.trap_313 = P.trap_e = dummy_trap
    elif role == "puls":
        P.trap_10410 = _trap(us(10.0), us(4.0), step)          # src/dsp_puls.jl:56
        P.trap_535 = P.trap_313 = P.trap_e = P.trap_10410      # unused columns: the same smooth trace (cheap pruning)
    else:
        P.trap_10410 = _trap(us(10.0), us(4.0), step)
        P.trap_535 = _trap(us(5.0), us(3.0), step)
        P.trap_313 = _trap(us(3.0), us(1.0), step)
        P.trap_e = _trap(trap_rt, trap_ft, step)
    P.trap_pickoff_ns = (trap_rt + trap_ft / 2).ns()
    P.cusp_pickoff_ns = (cfg.flt_length_cusp / 2).ns()
    P.zac_pickoff_ns = (cfg.flt_length_zac / 2).ns()
    for tr, name in ((P.t0_trap, "t0"), (P.t0inv_trap, "t0_inv"), (P.trap_10410, "10410"), (P.trap_535, "535"),
                     (P.trap_313, "313"), (P.trap_e, "trap")):
        if tr.navg < 1 or tr.navg2 < 1 or tr.ngap < 0 or tr.length > n:
            raise ValueError(f"trapezoidal filter {name} {tr.as_tuple()} does not fit {n} samples")

    # CUSP / ZAC: src/dsp_icpc.jl:87-90,98-99,167,174
    tau_off = us(10000000.0)
    if role == "wdw" or lite:
        for cz in (P.cusp, P.zac):
            cz.n_taps, cz.flat, cz.sigma, cz.tau, cz.beta = 4, 0, 1.0, 1.0, 1.0
    else:
        _fill_cuspzac(P.cusp, "cusp", cusp_rt, cusp_ft, tau_off, cfg.flt_length_cusp,
                      _ratio(cfg.flt_length_cusp, step), step, policy, builders)
        _fill_cuspzac(P.zac, "zac", zac_rt, zac_ft, tau_off, cfg.flt_length_zac,
                      _ratio(cfg.flt_length_zac, step), step, policy, builders)
    if P.cusp.n_taps > n or P.zac.n_taps > n:
        raise ValueError("CUSP/ZAC filter longer than the waveform")

    # currents: src/dsp_icpc.jl:181-186
    deg = int(cfg.sg_flt_degree)
    if lite:
        for k in range(3):
            _fill_sg(P.sg[k], step * 5.0, deg, step, policy, builders)
            P.cur_from[k], P.cur_until[k] = 0, n - P.sg[k].n_taps
        P.cur_from[3], P.cur_until[3] = 0, n - 1
    elif role == "pre":
        # :439 the in-trace / current-rise filter of the presummed waveform; the currents themselves (:431-435) are
        # taken from the windowed waveform
        for k in range(3):
            _fill_sg(P.sg[k], (sg_wl * float(presum_rate)) / 2.0, deg, step, policy, builders)
            P.cur_from[k], P.cur_until[k] = 0, n - P.sg[k].n_taps
        P.cur_from[3], P.cur_until[3] = 0, n - 1
    else:
        _fill_sg(P.sg[0], sg_wl, deg, step, policy, builders)
        _fill_sg(P.sg[1], ns(60.0), deg, step, policy, builders)
        _fill_sg(P.sg[2], ns(100.0), deg, step, policy, builders)
        for k in range(3):
            first_k = t_first + step * float(P.sg[k].offset)
            a, b = window(cfg.current_window, first_k, n - P.sg[k].n_taps + 1, f"current_window on sg[{k}]")
            P.cur_from[k], P.cur_until[k] = a, b
        a, b = window(cfg.current_window, t_first, n, "current_window")
        P.cur_from[3], P.cur_until[3] = a, b

    # get_intracePileUp: src/dsp_routines.jl:72-82
    P.intrace_nsigma = float(cfg.inTraceCut_std_threshold)
    P.intrace_min_n = _min_n(kw["intrace_mintot"], step)
    first_sg = t_first + step * float(P.sg[0].offset)
    n_sg = n - P.sg[0].n_taps + 1
    if role == "wdw" or lite:
        a, b = 0, min(n_sg - 1, 15)       # in-trace pile-up is evaluated on the presummed waveform (:440)
    else:
        a = _sub_over_step(cfg.bl_window[0] + first_sg, first_sg, step)   # leftendpoint + first(time), :75
        b = _sub_over_step(cfg.bl_window[1], first_sg, step)
    if not (0 <= a <= b <= n_sg - 1):
        raise AssertionError(f"in-trace sigma window {a + 1}:{b + 1} outside 1:{n_sg}")
    P.intrace_bl_from, P.intrace_bl_until = a, b

    P.cuspzac_direct = 1 if cuspzac_direct else 0
    return P


def resolve_compressed_params(cfg: DSPConfig, tau: Q, pars_filter: Optional[Dict[str, Any]] = None, *, presum_rate: int,
                               n_pre: int, t_first_pre: Q = ns(0.0), step_pre: Q, n_wdw: int, t_first_wdw: Q = ns(0.0),
                               step_wdw: Q = ns(16.0), policy: RddspPolicy = DEFAULT_POLICY, builders=None):
    """What `dsp_icpc_compressed` (src/dsp_icpc.jl:293-499) derives from (config, tau, pars_filter), the two time axes
    and the presum rate: (params of the presummed pass, params of the windowed pass, the four auxiliary windows
    auxbl1, auxbl2, auxpz1, auxpz2 as 0-based inclusive sample ranges of the presummed axis)."""
    P_pre = resolve_icpc_params(cfg, tau, pars_filter, n_samples=n_pre, t_first=t_first_pre, step=step_pre, policy=policy,
                                builders=builders, role="pre", presum_rate=presum_rate)
    P_wdw = resolve_icpc_params(cfg, tau, pars_filter, n_samples=n_wdw, t_first=t_first_wdw, step=step_wdw, policy=policy,
                                builders=builders, role="wdw", presum_rate=presum_rate)
    aux = []
    for name in ("auxbl1_window", "auxbl2_window", "auxpz1_window", "auxpz2_window"):
        w = getattr(cfg, name)
        a = _sub_over_step(w[0], t_first_pre, step_pre)
        b = _sub_over_step(w[1], t_first_pre, step_pre)
        # @assert firstindex(X) <= first(idxs) <= last(idxs) <= lastindex(X)   src/tailstats.jl:23-25
        if not (0 <= a <= b <= n_pre - 1):
            raise AssertionError(f"{name}: index range {a + 1}:{b + 1} outside 1:{n_pre}")
        aux.append((a, b))
    return P_pre, P_wdw, aux


def resolve_sweep_params(cfg: DSPConfig, tau: Q, *, n_samples: int = 8192, t_first: Q = ns(0.0),
                         step: Q = ns(16.0), builders=None, out_f64: bool = False, external_baseline: bool = False) -> _abi.SweepParams:
    """common part of dsp_trap_rt_optimization / dsp_trap_ft_optimization
    (src/dsp_filter_optimization.jl:102-133, 241-274)"""
    if builders is None:
        builders = LibBuilders()
    kw = cfg.kwargs_pars
    S = _abi.SweepParams()
    S.struct_size = C.sizeof(_abi.SweepParams)
    S.version = _abi.LGDSP_PARAMS_VERSION
    S.n_samples = int(n_samples)
    S.tx_min_n = _min_n(kw["tx_mintot"], step)
    S.t_first_ns = t_first.ns()
    S.dt_ns = step.ns()
    if external_baseline:
        # the waveform is shifted by a baseline the caller provides (windowed waveform of the compressed format, whose time
        # axis does not contain bl_window): the kernel's own window is a placeholder, its statistics are not read
        a, b = 0, min(int(n_samples) - 1, 15)
    else:
        a = _sub_over_step(cfg.bl_window[0], t_first, step)
        b = _sub_over_step(cfg.bl_window[1], t_first, step)
        if not (0 <= a <= b <= n_samples - 1):
            raise AssertionError(f"bl_window: index range {a + 1}:{b + 1} outside 1:{n_samples}")
    S.bl_from, S.bl_until = a, b
    RC = _ratio(tau, step)
    alpha = RC / (RC + 1.0)
    S.pz_km1 = 1.0 / alpha - 1.0
    _fill_dni(S.sig_dni, int(kw["sig_interpolation_order"]), kw["sig_interpolation_length"], step, builders)
    S.out_f64 = 1 if out_f64 else 0
    return S


def trap_variants(rts: Sequence[Q], fts: Sequence[Q], step: Q, *, mode: str, pickoff: Optional[Q] = None):
    """variant table for a sweep: mode "rt" -> fixed pick-off `pickoff` (enc_pickoff_trap),
    mode "ft" -> pick-off t50 + rt + ft/2.  Order: rt-major (for r in rts, for f in fts)."""
    out = (_abi.TrapVariant * (len(rts) * len(fts)))()
    i = 0
    for rt in rts:
        for ft in fts:
            out[i].trap = _trap(rt, ft, step)
            if mode == "rt":
                out[i].pickoff_ns = pickoff.ns()
                out[i].pickoff_mode = 0
            else:
                out[i].pickoff_ns = (rt + ft / 2).ns()
                out[i].pickoff_mode = 1
            i += 1
    return out


class SweepVariants:
    """ctypes array of lgdsp_sweep_variant plus the numpy coefficient arrays its pointers refer to (kept alive here)"""

    def __init__(self, n: int):
        self.array = (_abi.SweepVariant * n)()
        self._keep = []

    def __len__(self):
        return len(self.array)

    def set_coeffs(self, i: int, c: np.ndarray):
        c = np.ascontiguousarray(c, dtype=np.float64)
        self._keep.append(c)
        self.array[i].coeffs = c.ctypes.data_as(C.POINTER(C.c_double))
        self.array[i].n_taps = int(c.size)


def trap_sweep_variants(rts: Sequence[Q], fts: Sequence[Q], step: Q, *, mode: str, pickoff: Optional[Q] = None) -> SweepVariants:
    """the trapezoid sweeps as general variants (kind 0); same order and pick-offs as trap_variants"""
    tv = trap_variants(rts, fts, step, mode=mode, pickoff=pickoff)
    out = SweepVariants(len(tv))
    for i in range(len(tv)):
        out.array[i].kind = 0
        out.array[i].trap = tv[i].trap
        out.array[i].pickoff_ns = tv[i].pickoff_ns
        out.array[i].pickoff_mode = tv[i].pickoff_mode
    return out


def cuspzac_sweep_variants(cfg: DSPConfig, kind: str, rts: Sequence[Q], fts: Sequence[Q], step: Q, *, mode: str,
                           policy: RddspPolicy = DEFAULT_POLICY, builders=None) -> SweepVariants:
    """CUSP / ZAC sweeps (kind 1): for rt in rts, for ft in fts the filter `CUSPChargeFilter(rt, ft, 1e7 us, length, scale)`
    (src/dsp_filter_optimization.jl:173,221,316,366); mode "rt": fixed pick-off enc_pickoff_cusp/zac (:175,:223);
    mode "ft": t50 + flt_length/2 (:318,:368)"""
    if builders is None:
        builders = LibBuilders()
    length = cfg.flt_length_cusp if kind == "cusp" else cfg.flt_length_zac
    pick = cfg.enc_pickoff_cusp if kind == "cusp" else cfg.enc_pickoff_zac
    out = SweepVariants(len(rts) * len(fts))
    i = 0
    for rt in rts:
        for ft in fts:
            cz = _abi.CuspZac()
            _fill_cuspzac(cz, kind, rt, ft, us(10000000.0), length, _ratio(length, step), step, policy, builders)
            out.array[i].kind = 1
            out.set_coeffs(i, np.array(cz.coeffs[:cz.n_taps]))
            if mode == "rt":
                out.array[i].pickoff_ns = pick.ns()
                out.array[i].pickoff_mode = 0
            else:
                out.array[i].pickoff_ns = (length / 2).ns()
                out.array[i].pickoff_mode = 1
            i += 1
    return out


def sg_sweep_variants(cfg: DSPConfig, wls: Sequence[Q], *, n_samples: int, t_first: Q, step: Q,
                      policy: RddspPolicy = DEFAULT_POLICY, builders=None) -> SweepVariants:
    """Savitzky-Golay window-length sweep (kind 2): SavitzkyGolayFilter(wl, sg_flt_degree, 1) and get_wvf_maximum in
    current_window (src/dsp_filter_optimization.jl:430-433); the window is resolved on every filter's own trace axis"""
    if builders is None:
        builders = LibBuilders()
    out = SweepVariants(len(wls))
    for i, wl in enumerate(wls):
        sg = _abi.Sg()
        _fill_sg(sg, wl, int(cfg.sg_flt_degree), step, policy, builders)
        first_k = t_first + step * float(sg.offset)
        n_trace = n_samples - sg.n_taps + 1
        a = _sub_over_step(cfg.current_window[0], first_k, step)
        b = _sub_over_step(cfg.current_window[1], first_k, step)
        if not (0 <= a <= b <= n_trace - 1):
            raise AssertionError(f"current_window on sg(wl={wl}): index range {a + 1}:{b + 1} outside 1:{n_trace}")
        out.array[i].kind = 2
        out.set_coeffs(i, np.array(sg.h[:sg.n_taps]))
        out.array[i].sg_offset = sg.offset
        out.array[i].win_from, out.array[i].win_until = a, b
    return out


def params_summary(P: _abi.IcpcParams) -> Dict[str, Any]:
    """human-readable dump of the resolved sample-domain constants (compare with SURVEY.md Appendix A)"""
    return {
        "n_samples": P.n_samples, "dt_ns": P.dt_ns, "sat": (P.sat_low, P.sat_high),
        "bl": (P.bl_from, P.bl_until), "tail": (P.tail_from, P.tail_until),
        "RC": (1.0 / P.pz_km1) if P.pz_km1 != 0 else float("inf"), "t0_trap": P.t0_trap.as_tuple(), "t0_min_n": P.t0_min_n,
        "tx_min_n": P.tx_min_n, "intrace_min_n": P.intrace_min_n,
        "trap_10410": P.trap_10410.as_tuple(), "trap_535": P.trap_535.as_tuple(),
        "trap_313": P.trap_313.as_tuple(), "trap_e": P.trap_e.as_tuple(),
        "sig_dni": (P.sig_dni.degree, P.sig_dni.n_w), "int_dni": (P.int_dni.degree, P.int_dni.n_w),
        "cusp": (P.cusp.n_taps, P.cusp.flat, P.cusp.sigma), "zac": (P.zac.n_taps, P.zac.flat, P.zac.sigma),
        "sg": [(P.sg[k].n_taps, P.sg[k].offset) for k in range(3)],
        "cur": [(P.cur_from[k], P.cur_until[k]) for k in range(4)],
        "intrace_bl": (P.intrace_bl_from, P.intrace_bl_until),
    }
