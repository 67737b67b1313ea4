"""Pre-CMVN cepstra matrices [49][13] that no audio clip produces, shared by the CPU model test (test_cmvn_bound.py) and the
GPU test of the device implementation (test_gpu_parity.py): they stress the certified CMVN shortcut of the classify kernels
(csrc/kernels.cu cmvn_certified / cmvn_resolve) where its error bound is tightest."""
import numpy as np

EPS = np.float32(1.1920929e-07)


def pad_rows():
    """numpy::pad_1d_symmetric (numpy.hpp:479-541) for 49 rows padded by 50 on each side: source frame of every padded row"""
    rows, before, after = 49, 50, 50
    src = np.zeros(rows + before + after, np.int64)
    idx, up = 0, True
    for ix in range(before - 1, -1, -1):
        src[ix] = idx
        if idx == 0 and not up:
            up = True
        elif idx == rows - 1 and up:
            up = False
        elif up:
            idx += 1
        else:
            idx -= 1
    src[before:before + rows] = np.arange(rows)
    idx, up = rows - 1, False
    for ix in range(after):
        src[ix + before + rows] = idx
        if idx == 0 and not up:
            up = True
        elif idx == rows - 1 and up:
            up = False
        elif up:
            idx += 1
        else:
            idx -= 1
    return src


def adversarial_matrices(scale, seed=20261017):
    """the 72 hand-shaped matrices: huge mean/sigma ratios, near-constant columns, outliers, tiny and huge magnitudes, ties"""
    rng = np.random.default_rng(seed)
    scale = np.float32(scale)
    mats = []
    for sigma in (1e-3, 1.0, 30.0, 1e3):
        for ratio in (0.0, 10.0, 1e3, 1e5, -1e4):
            mats.append((rng.standard_normal((49, 13)) * sigma + ratio * sigma).astype(np.float32))
    for k in range(10):
        m = (rng.standard_normal((49, 13)) * 3).astype(np.float32)
        m[rng.integers(0, 49), :] += np.float32(10.0 ** rng.integers(1, 6))      # one outlier frame
        mats.append(m)
        mats.append((np.float32(7.25) + rng.standard_normal((49, 13)) * 1e-6).astype(np.float32))  # nearly constant
        mats.append((rng.standard_normal((49, 13)) * 10.0 ** rng.integers(-20, 15)).astype(np.float32))
        mats.append(np.where(rng.random((49, 13)) < 0.5, np.float32(1.0), np.float32(-1.0)).astype(np.float32) * np.float32(2.5))
        mats.append(np.round(rng.standard_normal((49, 13)) * 4).astype(np.float32) * scale)  # many exact-looking values
    mats.append(np.zeros((49, 13), np.float32))
    mats.append(np.full((49, 13), 3.0, np.float32))
    return mats


def special_value_matrices(seed=7):
    """denormal, huge, infinite and NaN columns / entries: every comparison of the bound must fail safely into the exact path"""
    rng = np.random.default_rng(seed)
    base = lambda: (rng.standard_normal((49, 13)) * 4).astype(np.float32)
    mats = []
    m = base(); m[:, 3] = np.float32(1e-42) * rng.integers(1, 100, 49); mats.append(m)          # denormal column
    m = base(); m[:, 0] = np.float32(1e-45); mats.append(m)                                      # constant smallest denormal
    m = base(); m[:, 5] = (rng.standard_normal(49) * 1e37).astype(np.float32); mats.append(m)    # sums of squares overflow float
    m = base(); m[:, 7] = np.float32(3e38) * np.where(rng.random(49) < 0.5, 1, -1); mats.append(m)
    m = base(); m[10, 2] = np.inf; mats.append(m)
    m = base(); m[20, 4] = -np.inf; m[21, 4] = np.inf; mats.append(m)
    m = base(); m[30, 6] = np.nan; mats.append(m)
    m = base(); m[:, 8] = np.nan; mats.append(m)
    m = base(); m[0, :] = np.float32(1e30); m[48, :] = np.float32(-1e30); mats.append(m)         # the mirrored edge frames
    m = base(); m[:, 9] = np.float32(-0.0); mats.append(m)
    m = base() * np.float32(1e-30); mats.append(m)
    m = base(); m[:, 11] = np.float32(65504.0); m[24, 11] = np.float32(65504.0 * (1 + 2 ** -23)); mats.append(m)  # variance of one ulp
    return mats


def near_boundary_matrices(n, scale, seed):
    """n random matrices in which ONE frame per column has been moved so that its normalised value t = f / scale sits on (or a few
    rounding errors from) a rounding boundary k + 1/2 of the int8 input quantisation -- where a shortcut that is only 'nearly'
    right would flip a bit.  t as a function of that one entry x is closed-form (the entry occurs `mult` times in its own padded
    window): solved by bisection in float64; the float32 rounding of x then scatters t around the boundary on the scale of the
    reference's own rounding errors.  Returns (matrices [n,49,13] float32, target frames [n,13])."""
    rng = np.random.default_rng(seed)
    src = pad_rows()
    mult = np.array([(src[r:r + 101] == r).sum() for r in range(49)], np.float64)
    scale = float(np.float32(scale))
    sig = 10.0 ** rng.uniform(-2, 2, size=(n, 1, 1))
    off = rng.standard_normal((n, 1, 13)) * sig * 10.0 ** rng.uniform(-1, 2.5, size=(n, 1, 1))
    F = (rng.standard_normal((n, 49, 13)) * sig + off).astype(np.float32)
    G = F.astype(np.float64)[:, src, :]
    PS = np.concatenate([np.zeros((n, 1, 13)), np.cumsum(G, axis=1)], axis=1)
    PQ = np.concatenate([np.zeros((n, 1, 13)), np.cumsum(G * G, axis=1)], axis=1)
    r = rng.integers(0, 49, size=(n, 13))
    ii, cc = np.arange(n)[:, None], np.arange(13)[None, :]
    x0 = F[ii, r, cc].astype(np.float64)
    m = mult[r]
    S0 = (PS[ii, r + 101, cc] - PS[ii, r, cc]) - m * x0
    Q0 = (PQ[ii, r + 101, cc] - PQ[ii, r, cc]) - m * x0 * x0
    del G, PS, PQ

    def t_of(x):
        S = S0 + m * x
        V = np.maximum(Q0 + m * x * x - S * S / 101.0, 0.0)
        return (x - S / 101.0) / (np.sqrt(V / 101.0) + float(EPS)) / scale

    t0 = t_of(x0)
    delta = rng.choice(np.array([0.0, 3e-8, -3e-8, 1e-7, -1e-7, 1e-6, -1e-6, 1e-5, -1e-5]), size=(n, 13))
    target = np.floor(t0) + 0.5 + delta * np.maximum(1.0, np.abs(t0))
    s = np.sqrt(np.maximum(Q0 / 100.0 - (S0 / 100.0) ** 2, 1e-30))
    lo, hi = x0 - 2.0 * s, x0 + 2.0 * s
    ok = (t_of(lo) < target) & (t_of(hi) > target)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        up = t_of(mid) < target
        lo = np.where(up, mid, lo)
        hi = np.where(up, hi, mid)
    x = np.where(ok, 0.5 * (lo + hi), x0).astype(np.float32)
    F[ii, r, cc] = x
    return F, r
