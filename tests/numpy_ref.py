"""Independent numpy restatement of the kernel-matrix fill and the flow, used ONLY to
cross-check the C oracle on small inputs (tests/test_oracle.py).  Written against the
same reference lines (CvoGPU.cu:477-593, :729-848) but vectorised per source row, so a
slip in one restatement shows up as a disagreement with the other."""
import numpy as np

f32 = np.float32


def transform(R, T, y):
    """R (3,3), T (3,) float32 = T_target_to_source blocks; returns y' = R^T (y - T) computed
    the way update_tf + transform_point_R_T do: Rinv = R^T, Tinv = -Rinv T, y' = Rinv y + Tinv
    with three-term sums c0 + (c1 + c2)."""
    R = np.asarray(R, f32)
    T = np.asarray(T, f32)
    Rinv = R.T.copy()
    neg = -Rinv
    Tinv = (neg[:, 0] * T[0] + (neg[:, 1] * T[1] + neg[:, 2] * T[2])).astype(f32)
    y = np.asarray(y, f32)
    out = np.empty_like(y)
    for i in range(3):
        r = Rinv[i, 0] * y[:, 0] + (Rinv[i, 1] * y[:, 1] + Rinv[i, 2] * y[:, 2])
        out[:, i] = r.astype(f32) + Tinv[i]
    return out


def fill_A(p, x, y_moved, ell, num_neighbors, fx=None, fy=None, lx=None, ly=None, gx=None, gy=None):
    """Returns (rows of (idx, val)) with the reference's truncation rule."""
    x = np.asarray(x, f32)
    ym = np.asarray(y_moved, f32)
    N, M = len(x), len(ym)
    sigma2 = f32(p.sigma) * f32(p.sigma)
    c2 = f32(p.c_ell) * f32(p.c_ell)
    c_sigma2 = f32(p.c_sigma) * f32(p.c_sigma)
    s_ell = f32(p.s_ell)
    s_sigma2 = f32(p.s_sigma) * f32(p.s_sigma)
    sp = f32(p.sp_thres)
    rows = []
    for i in range(N):
        pa = x[i]
        a_to_sensor = np.sqrt(f32(f32(pa[0] * pa[0]) + f32(pa[1] * pa[1])) + f32(pa[2] * pa[2]), dtype=f32)
        l = f32((np.float64(a_to_sensor) / 500.0 + 1.0) * np.float64(f32(ell)))
        keep = np.ones(M, bool)
        a = np.ones(M, f32)
        if p.is_using_geometric_type:
            ga = gx[i] if gx is not None else np.zeros(2, f32)
            gb = gy if gy is not None else np.zeros((M, 2), f32)
            n2a = f32(f32(0) + ga[0] * ga[0]) + ga[1] * ga[1]
            n2b = (f32(0) + gb[:, 0] * gb[:, 0]).astype(f32) + gb[:, 1] * gb[:, 1]
            dab = (f32(0) + ga[0] * gb[:, 0]).astype(f32) + ga[1] * gb[:, 1]
            with np.errstate(invalid="ignore", divide="ignore"):
                geo = (dab * dab).astype(f32) / (n2a * n2b).astype(f32)
            keep &= ~(geo.astype(np.float64) < 0.01)
        else:
            geo = np.ones(M, f32)
        k = np.ones(M, f32)
        if p.is_using_geometry:
            d2_thres = f32(-2.0 * np.float64(l) * np.float64(l) * np.float64(np.log(f32(sp / sigma2), dtype=f32)))
            d = ym - pa
            d2 = ((d[:, 0] * d[:, 0]).astype(f32) + (d[:, 1] * d[:, 1]).astype(f32)).astype(f32) + (d[:, 2] * d[:, 2]).astype(f32)
            keep &= d2 < d2_thres
            k = (np.float64(sigma2) * np.exp(-d2.astype(np.float64) / (2.0 * np.float64(l) * np.float64(l)))).astype(f32)
        ck = np.ones(M, f32)
        if p.is_using_intensity:
            d2_c_thres = f32(-2.0 * np.float64(c2) * np.float64(np.log(f32(sp / c_sigma2), dtype=f32)))
            F = 0 if fx is None else fx.shape[1]
            d2c = np.zeros(M, f32)
            for f in range(F):
                t = (fx[i, f] - fy[:, f]).astype(f32)
                d2c = (d2c + (t * t).astype(f32)).astype(f32)
            keep &= d2c < d2_c_thres
            ck = (np.float64(c_sigma2) * np.exp(-d2c.astype(np.float64) / (2.0 * np.float64(c2)))).astype(f32)
        sk = np.ones(M, f32)
        if p.is_using_semantics:
            d2_s_thres = f32(-2.0 * np.float64(s_ell) * np.float64(s_ell) * np.float64(np.log(f32(sp / s_sigma2), dtype=f32)))
            C = 0 if lx is None else lx.shape[1]
            d2s = np.zeros(M, f32)
            for c in range(C):
                t = (lx[i, c] - ly[:, c]).astype(f32)
                d2s = (d2s + (t * t).astype(f32)).astype(f32)
            keep &= d2s < d2_s_thres
            sk = (np.float64(s_sigma2) * np.exp(-d2s.astype(np.float64) / (2.0 * np.float64(s_ell) * np.float64(s_ell)))).astype(f32)
        a = (((ck * k).astype(f32) * sk).astype(f32) * geo).astype(f32)
        with np.errstate(invalid="ignore"):
            keep &= a > sp
        idx = np.nonzero(keep)[0][:num_neighbors]
        rows.append((idx.astype(np.int32), a[idx]))
    return rows


def flow(p, x, y_moved, rows):
    """compute_flow: returns (omega_sum, v_sum) float64 and the normalised float32 twist."""
    x = np.asarray(x, f32)
    ym = np.asarray(y_moved, f32)
    om_sum = np.zeros(3, np.float64)
    v_sum = np.zeros(3, np.float64)
    for i, (idx, val) in enumerate(rows):
        om = np.zeros(3, f32)
        vv = np.zeros(3, f32)
        px = x[i]
        for j, a in zip(idx, val):
            py = ym[j]
            cr = np.array([px[1] * py[2] - px[2] * py[1], px[2] * py[0] - px[0] * py[2],
                           px[0] * py[1] - px[1] * py[0]], f32)
            om = (om + (cr * a).astype(f32)).astype(f32)
            vv = (vv + ((py - px).astype(f32) * a).astype(f32)).astype(f32)
        om_sum += (om / f32(p.c)).astype(f32).astype(np.float64)
        v_sum += (vv / f32(p.d)).astype(f32).astype(np.float64)
    ov = np.concatenate([om_sum, v_sum]).astype(f32)
    n = np.linalg.norm(ov.astype(np.float64))
    tw = (ov / f32(n)).astype(f32) if n > 0 else ov
    return om_sum, v_sum, tw


def step_poly(p, x, y_moved, rows, omega, v, ell):
    """B, C, D, E of compute_step_size_poly_coeff (CvoGPU.cu:1001-1082) by a GENERIC route, in
    float64: the reference's hand-expanded beta..epsilon combinations are the Taylor coefficients
    of  g(t) = sum_ij A_ij exp(q_ij(t)),  q_ij(t) = -c_i (|D_ij + sum_k t^k xi^k z_j|^2 - |D_ij|^2),
    D_ij = x_i - y'_j, xi^k z = (xi_hat^k [y'_j; 1])_xyz (no factorials: CvoGPU.cu:953-998 feeds
    the raw powers), c_i = 1 / (2 l_i^2).  Here q is built by polynomial multiplication and exp(q)
    by the power-series recurrence e_n = (1/n) sum_k k q_k e_{n-k} - none of the reference's
    formulas is restated.  Returns (B, C, D, E) = sum_ij A_ij e_1..e_4."""
    x = np.asarray(x, np.float64)
    ym = np.asarray(y_moved, np.float64)
    om = np.asarray(omega, np.float64)
    xi = np.zeros((4, 4))
    xi[:3, :3] = [[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]]
    xi[:3, 3] = np.asarray(v, np.float64)
    yh = np.concatenate([ym, np.ones((len(ym), 1))], axis=1)
    powers, M = [], np.eye(4)
    for _ in range(4):
        M = M @ xi
        powers.append((yh @ M.T)[:, :3])  # xi^k z for every target point
    out = np.zeros(4)
    for i, (idx, val) in enumerate(rows):
        if len(idx) == 0:
            continue
        l = np.float64(ell)
        if p.is_using_range_ell:
            l = (np.linalg.norm(x[i]) / 500.0 + 1.0) * np.float64(f32(ell))  # CvoGPU.cu:86-90
        c = 1.0 / (2.0 * l * l)
        d0 = x[i] - ym[idx]                                   # (k, 3)
        coeffs = np.stack([d0] + [pw[idx] for pw in powers], axis=1)  # (k, 5, 3): t^0..t^4
        q = np.zeros((len(idx), 9))
        for a in range(5):
            for b in range(5):
                q[:, a + b] += (coeffs[:, a, :] * coeffs[:, b, :]).sum(1)
        q[:, 0] = 0.0                                          # - |D|^2
        q *= -c
        e = np.zeros((len(idx), 5))
        e[:, 0] = 1.0
        for n in range(1, 5):
            for k in range(1, n + 1):
                e[:, n] += k * q[:, k] * e[:, n - k]
            e[:, n] /= n
        out += (np.asarray(val, np.float64)[:, None] * e[:, 1:]).sum(0)
    return tuple(out)


def step_from_poly(p, B, C, D, E):
    """compute_step_size (CvoGPU.cu:1124-1158): smallest positive real root of g'(t) ~ 0, clamped."""
    roots = np.roots([4.0 * E, 3.0 * D, 2.0 * C, B])
    cand = [r.real for r in roots if r.real > 0 and abs(r.imag) < 1e-5]
    t = min(cand) if cand else np.inf
    return float(min(max(t, p.min_step), p.max_step)) if np.isfinite(t) else float(p.max_step)
