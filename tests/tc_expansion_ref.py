"""NumPy statement of the separable expansion behind the tcgen05 evaluation route (csrc/gp_eval_tc.cu).

Every functional the ScaSML correction needs of the surrogate (models/GP.py:630-687, 746-769) is
    sum_j  kernel_c(x, y_j) * polynomial(x, y_j),      kernel_c in {k, ky, kx}  (x-y, x-roll y, roll x-y),
and the Gaussian factorises, kernel_c = K_i K_j exp(a x . perm_c(y_j)).  Expanding the polynomial in per-point
monomials  F[f1] F[f2]  with per-centre coefficients gives
    value_out(x_i) = K_i sum_col F_i[f1(col)] F_i[f2(col)]  T[i, col],      T = P_c C_c,   P_c[i, j] = exp(a x_i . perm_c y_j),
i.e. two chained GEMMs (distance GEMM -> exp -> coefficient GEMM), which is what the kernel runs on the tensor cores.
This module builds the column table and the coefficient matrices in float64 exactly as csrc/gp_eval_tc.cu does (same
order), so the algebra is tested on the CPU against the oracle's closed forms (tests/test_tc_expansion.py), and it can
emulate the f16 hi/lo operand splits to predict the route's accuracy.  TEST INFRASTRUCTURE, not shipped.
"""
import numpy as np

MC = 5
MODE_U, MODE_UG, MODE_PDE = 0, 1, 2
CLS_K, CLS_KY, CLS_KX = 0, 1, 2
OUT_U, OUT_G, OUT_L, OUT_T = 0, 1, 2, 3
# per-point features
F_ONE, F_SX, F_XT, F_X0, F_SXR, F_P2, F_R2, F_XI, F_XR = 0, 1, 2, 3, 4, 5, 6, 7, 12
NFEAT = 17


def point_features(X, idx, d):
    X = np.asarray(X, dtype=np.float64)
    F = np.empty((len(X), NFEAT))
    sx = X[:, :d].sum(1)
    F[:, F_ONE] = 1.0
    F[:, F_SX] = sx
    F[:, F_XT] = X[:, d]
    F[:, F_X0] = X[:, 0]
    F[:, F_SXR] = sx - X[:, 0] + X[:, d]
    F[:, F_P2] = (X[:, idx] ** 2).sum(1)
    F[:, F_R2] = (X[:, idx + 1] ** 2).sum(1)
    F[:, F_XI:F_XI + MC] = X[:, idx]
    F[:, F_XR:F_XR + MC] = X[:, idx + 1]
    return F


def columns(mode, cls):
    """[(out, f1, f2, coefficient id, m, n)] in kernel order for one (mode, kernel class)."""
    cols = []
    H = [F_ONE, F_P2] + [F_XI + m for m in range(MC)]          # monomials of h (ky class)
    Mx = [F_ONE, F_R2] + [F_XR + m for m in range(MC)]         # monomials of MHx (kx class)
    if cls == CLS_K:
        cols += [(OUT_U, F_ONE, F_ONE, "U0", 0, 0), (OUT_U, F_XT, F_ONE, "U1", 0, 0), (OUT_U, F_SX, F_ONE, "U2", 0, 0)]
        if mode >= MODE_UG:
            cols += [(OUT_G, F_ONE, F_ONE, "G0", 0, 0), (OUT_G, F_SX, F_ONE, "GSX", 0, 0), (OUT_G, F_XT, F_ONE, "GXT", 0, 0),
                     (OUT_G, F_SX, F_XT, "GSXXT", 0, 0), (OUT_G, F_SX, F_SX, "GSX2", 0, 0)]
        if mode == MODE_PDE:
            cols += [(OUT_T, F_ONE, F_ONE, "T0", 0, 0), (OUT_T, F_XT, F_ONE, "TXT", 0, 0), (OUT_T, F_SX, F_ONE, "TSX", 0, 0),
                     (OUT_T, F_XT, F_XT, "TXT2", 0, 0), (OUT_T, F_SX, F_XT, "TSXXT", 0, 0)]
            cols += [(OUT_L, F_ONE, F_ONE, "L1", 0, 0), (OUT_L, F_R2, F_ONE, "LR2", 0, 0), (OUT_L, F_R2, F_R2, "LR22", 0, 0)]
            cols += [(OUT_L, F_XR + m, F_ONE, "LX", m, 0) for m in range(MC)]
            cols += [(OUT_L, F_XR + m, F_R2, "LXR2", m, 0) for m in range(MC)]
            cols += [(OUT_L, F_XR + m, F_XR + n, "LXX", m, n) for m in range(MC) for n in range(m, MC)]
    elif cls == CLS_KY:
        cols += [(OUT_U, f, F_ONE, "H", i, 0) for i, f in enumerate(H)]
        if mode >= MODE_UG:
            cols += [(OUT_G, f, F_SX, "HSX", i, 0) for i, f in enumerate(H)]
            cols += [(OUT_G, f, F_ONE, "HG", i, 0) for i, f in enumerate(H)]
        if mode == MODE_PDE:
            cols += [(OUT_T, f, F_XT, "HXT", i, 0) for i, f in enumerate(H)]
            cols += [(OUT_T, f, F_ONE, "HT", i, 0) for i, f in enumerate(H)]
    else:
        if mode == MODE_PDE:
            G = [F_ONE, F_X0, F_SXR]
            cols += [(OUT_L, f, g, "MX", i, j) for i, f in enumerate(Mx) for j, g in enumerate(G)]
    return cols


def centre_coefficient(name, m, n, a, d, A1, A3, A4, A5, y, idx):
    """Coefficient `name` of one centre y (alphas already multiplied by K_j); mirrors coef_of() in gp_eval_tc.cu."""
    sy, yt, y0 = y[:d].sum(), y[d], y[0]
    syr = sy - y0 + yt
    yI, yr = y[idx], y[idx + 1]
    Q1, Q2, T1, T2 = yI.sum(), (yI ** 2).sum(), yr.sum(), (yr ** 2).sum()
    a2, a3, a4 = a * a, a ** 3, a ** 4
    dd = float(d)
    if name == "U0": return A1 - a * A4 * yt - a * A5 * sy
    if name == "U1": return a * A4
    if name == "U2": return a * A5
    if name == "G0": return a * A1 * sy - a2 * A4 * yt * sy + a * dd * A5 - a2 * A5 * sy * sy
    if name == "GSX": return -a * A1 + a2 * A4 * yt + 2 * a2 * A5 * sy
    if name == "GXT": return a2 * A4 * sy
    if name == "GSXXT": return -a2 * A4
    if name == "GSX2": return -a2 * A5
    if name == "T0": return a * A1 * yt + a * A4 - a2 * A4 * yt * yt - a2 * A5 * yt * sy
    if name == "TXT": return -a * A1 + 2 * a2 * A4 * yt + a2 * A5 * sy
    if name == "TSX": return a2 * A5 * yt
    if name == "TXT2": return -a2 * A4
    if name == "TSXXT": return -a2 * A5
    LW = A3 * dd * dd / (MC * MC)
    if name == "L1": return LW * (a4 * T2 * T2 - 14 * a3 * T2 + 35 * a2)
    if name == "LR2": return LW * (2 * a4 * T2 - 14 * a3)
    if name == "LR22": return LW * a4
    if name == "LX": return LW * (-4 * a4 * T2 + 28 * a3) * yr[m]
    if name == "LXR2": return -4 * LW * a4 * yr[m]
    if name == "LXX": return LW * 4 * a4 * yr[m] * yr[n] * (1.0 if m == n else 2.0)
    w3 = A3 * dd
    hc = [w3 * (a2 / MC * T2 - a), w3 * a2 / MC] + [-2 * w3 * a2 / MC * yr[i] for i in range(MC)]
    if name == "H": return hc[m]
    if name == "HSX": return -a * hc[m]
    if name == "HG":
        extra = -w3 * 2 * a2 / MC * T1 if m == 0 else (w3 * 2 * a2 / MC if m >= 2 else 0.0)
        return a * syr * hc[m] + extra
    if name == "HXT": return -a * hc[m]
    if name == "HT": return a * y0 * hc[m]
    if name == "MX":
        mc = [a2 / MC * Q2 - a, a2 / MC] + [-2 * a2 / MC * yI[i] for i in range(MC)]
        pc = [dd * (A1 - a * A4 * yt - a * A5 * sy), dd * a * A4, dd * a * A5]
        v = mc[m] * pc[n]
        if n == 0 and m == 0: v += dd * 2 * a2 / MC * A5 * Q1
        if n == 0 and m >= 2: v += -dd * 2 * a2 / MC * A5
        return v
    raise KeyError(name)


def perm_centres(Y, cls):
    if cls == CLS_K:
        return Y
    if cls == CLS_KY:                                   # roll(y) = (y_1, ..., y_d, y_0)
        return np.concatenate([Y[:, 1:], Y[:, :1]], axis=1)
    return np.concatenate([Y[:, -1:], Y[:, :-1]], axis=1)   # roll(x) . y = x . rollinv(y)


def build(mode, gp_o):
    """-> list over classes of (cols, C [Ncentres, ncol], Yperm [Ncentres, D]); centres = [domain | boundary]."""
    a, d, idx = gp_o.a, gp_o.d, gp_o.idx_set
    a1, a2b, a3, a4, a5 = gp_o._alpha_blocks()
    YD, YB = gp_o.x_t_domain, gp_o.x_t_boundary
    Y = np.concatenate([YD, YB], axis=0)
    nD = len(YD)
    Kj = np.exp(-0.5 * a * (Y * Y).sum(1))
    z = np.zeros(len(YB))
    A1 = np.concatenate([a1, a2b]) * Kj
    A3 = np.concatenate([a3, z]) * Kj
    A4 = np.concatenate([a4, z]) * Kj
    A5 = np.concatenate([a5, z]) * Kj
    out = []
    for cls in (CLS_K, CLS_KY, CLS_KX):
        cols = columns(mode, cls)
        if not cols:
            continue
        n = nD if cls == CLS_KY else len(Y)                 # the ky class has no boundary centres (w3 = 0 there)
        Cm = np.empty((n, len(cols)))
        for j in range(n):
            for c, (_, _, _, name, m, nn) in enumerate(cols):
                Cm[j, c] = centre_coefficient(name, m, nn, a, d, A1[j], A3[j], A4[j], A5[j], Y[j], idx)
        out.append((cls, cols, Cm, perm_centres(Y[:n], cls)))
    return out


def _split_f16(V):
    hi = V.astype(np.float16)
    lo = (V - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def evaluate(mode, gp_o, X, emulate=False, p_shift=6):
    """-> dict out -> values [R].  emulate=True applies the kernel's operand roundings: a' x = hi + lo (f16), P and the
    column-scaled coefficients split hi + lo (f16), products hh + hl + lh, float32 accumulators."""
    X = np.asarray(X, dtype=np.float64)
    a, d, idx = gp_o.a, gp_o.d, gp_o.idx_set
    F = point_features(X, idx, d)
    Ki = np.exp(-0.5 * a * (X * X).sum(1))
    res = {o: np.zeros(len(X)) for o in (OUT_U, OUT_G, OUT_L, OUT_T)}
    asc = a * 1.4426950408889634
    for cls, cols, Cm, Yp in build(mode, gp_o):
        if emulate:
            sv = (asc * X).astype(np.float32).astype(np.float64)
            xh, xl = _split_f16(sv)
            S = ((xl @ Yp.T).astype(np.float32) + (xh @ Yp.T).astype(np.float32)).astype(np.float32)
            P = np.exp2(S.astype(np.float64) + p_shift).astype(np.float32).astype(np.float64)
            ph, pl = _split_f16(P)
            cmax = np.abs(Cm).max(0)
            sc = np.where(cmax > 0, 2.0 ** (13 - np.ceil(np.log2(np.maximum(cmax, 1e-300)))), 1.0)
            ch, cl = _split_f16(Cm * sc)
            T = ((ph @ ch).astype(np.float32) + (ph @ cl).astype(np.float32) + (pl @ ch).astype(np.float32)).astype(np.float64)
            T = T / sc / 2.0 ** p_shift
        else:
            T = np.exp(a * (X @ Yp.T)) @ Cm
        for c, (o, f1, f2, *_r) in enumerate(cols):
            res[o] += F[:, f1] * F[:, f2] * T[:, c]
    return {o: Ki * v for o, v in res.items()}
