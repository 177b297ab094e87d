"""numpy prototype of the structured CUSP/ZAC evaluation (sliding exponential + polynomial windows on
d[j] = y[j] - r*y[j-1]); validates the closed forms and the step recurrences against the direct FIR."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import legenddsp.jl_b200 as L
from oracle import oracle as O

def descr(sigma, F, tau, Ltap, beta, zac):
    lt = (Ltap - F) // 2
    Rn = Ltap - lt - F - 1
    h = lt / 2.0
    a = 1.0 / np.sinh(lt / sigma)
    k = np.arange(Ltap)
    cusp = np.where(k < lt, np.sinh(k / sigma) * a, np.where(k <= lt + F, 1.0, np.sinh((Ltap - k) / sigma) * a))
    par = np.where(k < lt, (k - h) ** 2 - h * h, np.where(k <= lt + F, 0.0, (Ltap - k - h) ** 2 - h * h))
    B = -(cusp.sum() / par.sum()) if zac else 0.0
    c = cusp + B * par
    return dict(L=Ltap, F=F, lt=lt, Rn=Rn, h=h, a=a, B=B, sigma=sigma, r=np.exp(-1.0 / tau), g=beta / Ltap, c=c)

def structured(y, D):
    n = len(y); Lt, F, lt, Rn, h, a, B, sg, r, g = (D[k] for k in ("L","F","lt","Rn","h","a","B","sigma","r","g"))
    rho = np.exp(-1.0 / sg)
    ypad = np.concatenate([[0.0], y])
    d = ypad[1:] - r * ypad[:-1]           # d[j], j = 0..n-1 (y[-1] = 0)
    # prefix structures (index j -> value at sample j); use index shift +1 so that X[j+1] = prefix at j, X[0] = 0
    Pm = np.zeros(n + 1)                   # causal decayed: Pm[j+1] = sum_{i<=j} rho^(j-i) d[i]
    for j in range(n): Pm[j + 1] = rho * Pm[j] + d[j]
    Pp = np.zeros(n + 2)                   # anti-causal: Pp[j] = sum_{i>=j} rho^(i-j) d[i], Pp[n] = 0
    for j in range(n - 1, -1, -1): Pp[j] = rho * Pp[j + 1] + d[j]
    idx = np.arange(n, dtype=float)
    D0 = np.concatenate([[0.0], np.cumsum(d)])
    D1 = np.concatenate([[0.0], np.cumsum(idx * d)])
    D2 = np.concatenate([[0.0], np.cumsum(idx * idx * d)])
    pm = lambda j: Pm[j + 1]               # j >= -1
    pp = lambda j: Pp[j]                   # j in 0..n
    d0 = lambda j: D0[j + 1]; d1 = lambda j: D1[j + 1]; d2 = lambda j: D2[j + 1]
    nout = n - Lt + 1
    out = np.zeros(nout)
    for nn in range(nout):
        m = nn + Lt - 1
        EmL = pm(m) - rho ** lt * pm(m - lt)
        EpL = rho ** (-(lt - 1)) * (pp(m - lt + 1) - rho ** lt * pp(m + 1))
        EpR = rho ** (-Rn) * (pm(m - Lt + Rn) - rho ** Rn * pm(m - Lt))
        EmR = rho * (pp(m - Lt + 1) - rho ** Rn * pp(m - Lt + Rn + 1))
        # left flank polynomial: k in [0, lt): i in (m-lt, m]
        a0 = d0(m) - d0(m - lt); a1 = d1(m) - d1(m - lt); a2 = d2(m) - d2(m - lt)
        W1 = m * a0 - a1; W2 = m * m * a0 - 2 * m * a1 + a2
        # flat: i in [m-lt-F, m-lt]
        W0F = d0(m - lt) - d0(m - lt - F - 1)
        # right flank: k' = i-(m-L) in [1, Rn], i in [m-L+1, m-L+Rn]
        b0 = d0(m - Lt + Rn) - d0(m - Lt); b1 = d1(m - Lt + Rn) - d1(m - Lt); b2 = d2(m - Lt + Rn) - d2(m - Lt)
        q = m - Lt
        V1 = b1 - q * b0; V2 = b2 - 2 * q * b1 + q * q * b0
        Dm = 0.5 * a * (EpL - EmL + EpR - EmR) + B * (W2 - 2 * h * W1 + V2 - 2 * h * V1) + W0F
        ylast = y[m - Lt] if m - Lt >= 0 else 0.0
        out[nn] = g * (Dm + r * D["c"][Lt - 1] * ylast)
    return out

if __name__ == "__main__":
    P = L.resolve_icpc_params(L.example_config(), L.us(500.0), builders=O.OracleBuilders())
    wf = L.synth.generate_host(2, first_event=5)
    for ev in range(2):
        x = wf[ev].astype(float)
        bl = O.signalstats(x, 0.0, 16.0, P.bl_from, P.bl_until)
        y = O.invcr(x - bl["mean"], P.pz_km1)
        n = 2048 + 512
        y = y[2400:2400 + n]  # shorter for the slow python loops; includes the pulse
        for zac in (0, 1):
            cz = P.zac if zac else P.cusp
            Lt, F = 601, 40
            sigma = 80.5
            co = (O.OracleBuilders().zac_coeffs if zac else O.OracleBuilders().cusp_coeffs)(sigma, F, cz.tau, Lt, float(Lt))
            ref = O.fir_valid(y, co)
            got = structured(y, descr(sigma, F, cz.tau, Lt, float(Lt), zac))
            print("event", ev, "zac" if zac else "cusp", "max|ref|", np.abs(ref).max(), "max abs err", np.abs(ref - got).max())


def chunked(y, D, CH=32):
    """emulates the CUDA plan: decimated prefix tables at chunk-relative offsets + in-chunk recurrences"""
    n = len(y); Lt, F, lt, Rn, h, a, B, sg, r, g = (D[k] for k in ("L","F","lt","Rn","h","a","B","sigma","r","g"))
    rho = np.exp(-1.0 / sg); cA = 0.5 * a
    TT = np.concatenate([[0.0], np.cumsum(y)])          # TT[i] = sum_{k<i} y[k]
    def tt(i): return TT[i] if i >= 0 else 0.0
    def dd(j): return ((tt(j + 1) - tt(j)) - r * (tt(j) - tt(j - 1))) if 0 <= j < n else 0.0  # y[j] - r*y[j-1], differences exact
    nch = (n + CH - 1) // CH
    # full-resolution prefixes only to fill the decimated tables (the kernel gets them from scans)
    d = np.array([dd(j) for j in range(n)])
    Pm = np.zeros(n); acc = 0.0
    for j in range(n): acc = rho * acc + d[j]; Pm[j] = acc
    Pp = np.zeros(n + 1); acc = 0.0
    for j in range(n - 1, -1, -1): acc = rho * acc + d[j]; Pp[j] = acc
    idx = np.arange(n, dtype=float)
    D1 = np.cumsum(idx * d); D2 = np.cumsum(idx * idx * d)
    oc = [0, (-lt) % CH, (-lt - F - 1) % CH, (-Lt) % CH]
    oa = [(1 - lt) % CH, 1 % CH, (1 - Lt) % CH, (-lt - F) % CH]
    Tc = {q: {nm: np.array([arr[c * CH + oc[q]] if c * CH + oc[q] < n else np.nan for c in range(nch)])
              for nm, arr in (("Pm", Pm), ("D1", D1), ("D2", D2))} for q in range(4)}
    Ta = {q: np.array([Pp[c * CH + oa[q]] if c * CH + oa[q] <= n else np.nan for c in range(nch + 1)]) for q in range(4)}
    def look_c(q, nm, pos):
        if pos < 0: return 0.0
        assert pos % CH == oc[q]
        return Tc[q][nm][pos // CH]
    def look_a(q, pos):
        if pos < 0: return rho ** (-pos) * Pp[0]
        if pos >= n: return 0.0
        assert pos % CH == oa[q]
        return Ta[q][pos // CH]
    def d0(j): return (tt(j + 1) - r * tt(j)) if j >= 0 else 0.0
    nout = n - Lt + 1
    out = np.full(nout, np.nan)
    for t in range(nch):
        m0 = t * CH
        if m0 + CH - 1 < Lt - 1: continue
        m = m0
        EmL = cA * (look_c(0, "Pm", m) - rho ** lt * look_c(1, "Pm", m - lt))
        EpL = cA * rho ** (-(lt - 1)) * (look_a(0, m - lt + 1) - rho ** lt * look_a(1, m + 1))
        EpR = cA * rho ** (-Rn) * (look_c(2, "Pm", m - Lt + Rn) - rho ** Rn * look_c(3, "Pm", m - Lt))
        EmR = cA * rho * (look_a(2, m - Lt + 1) - rho ** Rn * look_a(3, m - Lt + Rn + 1))
        a0 = d0(m) - d0(m - lt); a1 = look_c(0, "D1", m) - look_c(1, "D1", m - lt); a2 = look_c(0, "D2", m) - look_c(1, "D2", m - lt)
        W0L = a0; W1L = m * a0 - a1; W2L = m * m * a0 - 2 * m * a1 + a2
        W0F = d0(m - lt) - d0(m - lt - F - 1)
        b0 = d0(m - Lt + Rn) - d0(m - Lt)
        b1 = look_c(2, "D1", m - Lt + Rn) - look_c(3, "D1", m - Lt); b2 = look_c(2, "D2", m - Lt + Rn) - look_c(3, "D2", m - Lt)
        q_ = m - Lt
        V0 = b0; V1 = b1 - q_ * b0; V2 = b2 - 2 * q_ * b1 + q_ * q_ * b0
        for k in range(CH):
            m = m0 + k
            if m >= n: break
            if m >= Lt - 1:
                Dm = (EpL - EmL + EpR - EmR) + B * (W2L - 2 * h * W1L + V2 - 2 * h * V1) + W0F
                ylast = (tt(m - Lt + 1) - tt(m - Lt)) if m - Lt >= 0 else 0.0
                out[m - Lt + 1] = g * (Dm + r * D["c"][Lt - 1] * ylast)
            s0, s1, s2, s3 = dd(m + 1), dd(m + 1 - lt), dd(m - lt - F), dd(m + 1 - Lt)
            EmL, EpL = cA * s0 + rho * EmL - cA * rho ** lt * s1, cA * s0 + EpL / rho - cA * rho ** (-lt) * s1
            W2L = W2L + 2 * W1L + W0L - lt * lt * s1
            W1L = W1L + W0L - lt * s1
            W0L = W0L + s0 - s1
            W0F = W0F + s1 - s2
            V2 = V2 - 2 * V1 + V0 + Rn * Rn * s2
            V1 = V1 - V0 + Rn * s2
            V0 = V0 + s2 - s3
            EpR = rho * EpR - cA * s3 + cA * rho ** (-Rn) * s2
            EmR = EmR / rho - cA * s3 + cA * rho ** Rn * s2
    return out


if __name__ == "__main__":
    wf = L.synth.generate_host(1, first_event=6)
    x = wf[0].astype(float)
    bl = O.signalstats(x, 0.0, 16.0, P.bl_from, P.bl_until)
    y = O.invcr(x - bl["mean"], P.pz_km1)[2400:2400 + 2560]
    for zac in (0, 1):
        Lt, F, sigma = 601, 40, 80.5
        co = (O.OracleBuilders().zac_coeffs if zac else O.OracleBuilders().cusp_coeffs)(sigma, F, P.cusp.tau, Lt, float(Lt))
        ref = O.fir_valid(y, co)
        got = chunked(y, descr(sigma, F, P.cusp.tau, Lt, float(Lt), zac))
        print("chunked", "zac" if zac else "cusp", "max|ref|", np.abs(ref).max(), "max abs err", np.nanmax(np.abs(ref - got)), "nan", np.isnan(got).sum())
