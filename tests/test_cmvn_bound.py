"""Host model of the certified CMVN shortcut (kernels.cu: cmvn_certified / cmvn_resolve level 2, derivation in DESIGN.md 4a),
checked against the plain-C oracle on CPU: whenever the bound certifies a rounding decision, the oracle's quantised feature
must be that integer, and the oracle's t = f / scale must lie within the bound B of the double-precision value.  Float32
arithmetic of the kernel is mirrored with numpy float32 (the MUFU approximations are replaced by correctly rounded ops; their
2^-21 allowance is part of B's 96u)."""
import numpy as np
import pytest

from oracle_lib import PortOracle

U = np.float32(5.9604645e-8)
EPS = np.float32(1.1920929e-07)


def pad_rows(rows=49, pad=50):
    src, idx, up = [0] * (rows + 2 * pad), 0, True
    for ix in range(pad - 1, -1, -1):
        src[ix] = idx
        if idx == 0 and not up:
            up = True
        elif idx == rows - 1 and up:
            up = False
        elif up:
            idx += 1
        else:
            idx -= 1
    for ix in range(rows):
        src[pad + ix] = ix
    idx, up = rows - 1, False
    for ix in range(pad):
        src[ix + pad + rows] = idx
        if idx == 0 and not up:
            up = True
        elif idx == rows - 1 and up:
            up = False
        elif up:
            idx += 1
        else:
            idx -= 1
    return np.array(src)


def certify(S, Q, Qall, x, inv_scale, em_term=True):
    """float32 part of cmvn_certified for arrays of windows: returns (ok, k, t_c, B)"""
    f32 = np.float32
    with np.errstate(all="ignore"):
        M = S * (1.0 / 101.0)
        V = Q - S * M
        xm = (x.astype(np.float64) - M).astype(f32)
        var = (V * (1.0 / 101.0)).astype(f32)
        qa = Qall.astype(f32)
        sig = np.sqrt(var)
        r = f32(1.0) / (sig + EPS)
        em = np.sqrt(qa)
        rv = f32(1.0) / (var * f32(101.0))
        ris = r * f32(inv_scale)
        tc = xm * ris
        relv = (f32(3.9e-11) * qa) * rv
        c_em = f32(1.0001) * U * f32(10.04987562) if em_term else f32(0.0)
        B = f32(1.02) * (np.abs(tc) * (f32(0.505) * relv + f32(96.0) * U) + (c_em * em) * ris) + f32(1e-30)
        k = np.rint(tc)
        dist = f32(0.5) - np.abs(tc - k)
        ok = (dist > B) & (relv < f32(9.765625e-4)) & (var > f32(1e-12)) & (np.abs(tc) < f32(1048576.0))
    return ok, k, tc, B


@pytest.mark.parametrize("name", ["l476", "gsc12"])
def test_certified_decisions_match_the_oracle(name, synth):
    port = PortOracle(name)
    clips = np.concatenate([synth.synth_clips(48, first_clip=31337, seed=0xABCD), np.stack(list(synth.special_clips().values()))])
    feats, taps = port.mfcc_i16(clips, taps=True)
    _, tens = port.run_inference(feats, want_tensors=True)
    want_q = np.stack([t[0] for t in tens]).view(np.int8).reshape(len(clips), 49, 13)
    # quantisation parameters of tensor 0, recovered from the oracle's own input tensor: q = (int8)(round(f / scale) + zp)
    src = pad_rows()
    scale, zp = _input_quant(name)
    inv_scale = np.float32(1.0 / float(scale))
    certified = total = 0
    worst = 0.0
    for ci in range(len(clips)):
        F32 = taps[ci]["mfcc"]
        G = F32.astype(np.float64)[src]
        PS = np.concatenate([np.zeros((1, 13)), np.cumsum(G, axis=0)])
        PQ = np.concatenate([np.zeros((1, 13)), np.cumsum(G * G, axis=0)])
        S, Q = PS[101:150] - PS[0:49], PQ[101:150] - PQ[0:49]
        f_ref = feats[ci].reshape(49, 13)
        with np.errstate(all="ignore"):
            t_ref = (f_ref / np.float32(scale)).astype(np.float32)
        for level in (1, 2):
            if level == 1:
                ok, k, tc, B = certify(S, Q, Q + 4.0 * np.max(G * G, axis=0)[None, :], F32, inv_scale)  # Q_all >= Q: any value above is admissible
            else:
                # level 2: the reference's own float mean (sequential float32 sum of the window)
                mean = np.zeros((49, 13), np.float32)
                for r in range(49):
                    acc = np.zeros(13, np.float32)
                    for w in range(101):
                        acc = (acc + G[r + w].astype(np.float32)).astype(np.float32)
                    mean[r] = acc / np.float32(101.0)
                M = S / 101.0
                dm = mean.astype(np.float64) - M
                V2 = (Q - S * M) + 101.0 * dm * dm
                S2 = mean.astype(np.float64) * 101.0  # makes certify()'s M equal the reference mean and its V equal V2
                ok, k, tc, B = certify(S2, V2 + S2 * (S2 / 101.0), Q, F32, inv_scale, em_term=False)
            with np.errstate(all="ignore"):
                k_ref = np.where(t_ref >= 0, np.floor(t_ref.astype(np.float64) + 0.5), np.ceil(t_ref.astype(np.float64) - 0.5))
                q_from_k = ((k.astype(np.int64) + zp) & 0xFF).astype(np.uint8).view(np.int8)
            assert not np.any(ok & (k_ref != k)), f"clip {ci} level {level}: a certified rounding decision differs from the oracle"
            assert np.array_equal(q_from_k[ok], want_q[ci][ok])
            with np.errstate(all="ignore"):
                ratio = np.abs(t_ref.astype(np.float64) - tc.astype(np.float64))[ok] / B[ok].astype(np.float64)
            if ratio.size:
                worst = max(worst, float(ratio.max()))
            if level == 1:
                certified += int(ok.sum())
                total += ok.size
    assert worst < 0.5, f"the oracle came within {worst:.2f} B of the bound: the derivation's constants need another look"
    # the shortcut must actually certify nearly everything on non-degenerate clips (40 of the 48 synthetic clips are)
    assert certified > 0.8 * total


def _input_quant(name):
    """(scale, zero_point) of the model's int8 input tensor, read from the oracle: quantise two probe features"""
    port = PortOracle(name)
    probe = np.zeros((1, port.n_features), np.float32)
    _, t0 = port.run_inference(probe, want_tensors=True)
    zp = int(t0[0][0].view(np.int8)[0])
    # grow f until the quantised value is 40 steps above the zero point (far from the int8 wrap), then bisect the 39.5 boundary
    def steps(f):
        probe[0, 0] = f
        _, t = port.run_inference(probe, want_tensors=True)
        return int(t[0][0].view(np.int8)[0]) - zp

    hi = 1e-6
    while steps(hi) < 40:
        hi *= 1.5
    lo = 0.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if steps(mid) >= 40:
            hi = mid
        else:
            lo = mid
    return np.float32(hi / 39.5), zp


def reference_cmvn_f32(F32):
    """numpy float32 emulation of processing::cmvnw + mean_axis0 + std_axis0 (processing.hpp:326-389, numpy.hpp:746-836) for a
    [49][13] cepstra matrix, vectorised over the 637 chains: every chain still performs the reference's operations in order"""
    f32, f64 = np.float32, np.float64
    G = F32[pad_rows()]
    rows = np.arange(49)
    s = np.zeros((49, 13), f32)
    for w in range(101):
        s = (s + G[rows + w]).astype(f32)
    mean = (s / f32(101.0)).astype(f32)
    sd = np.zeros((49, 13), f32)
    for w in range(101):
        d = (G[rows + w] - mean).astype(f32).astype(f64)
        sd = (sd.astype(f64) + d * d).astype(f32)
    std = np.sqrt((sd / f32(101.0)).astype(f32)).astype(f32)
    return ((F32 - mean).astype(f32) / (std + EPS).astype(f32)).astype(f32)


def test_emulation_is_the_oracle(synth):
    """pins reference_cmvn_f32 (used for the adversarial matrices below) to the plain-C oracle on real clips"""
    port = PortOracle("l476")
    clips = np.concatenate([synth.synth_clips(12, first_clip=99), np.stack(list(synth.special_clips().values()))])
    feats, taps = port.mfcc_i16(clips, taps=True)
    for ci in range(len(clips)):
        got = reference_cmvn_f32(taps[ci]["mfcc"])
        want = feats[ci].reshape(49, 13)
        assert np.all((got == want) | (np.isnan(got) & np.isnan(want))), f"clip {ci}"


def test_bound_on_adversarial_matrices():
    """cepstra matrices no audio clip produces: huge mean/sigma ratios, near-constant columns, outliers, tiny and huge magnitudes,
    exact ties.  A certified decision must always equal the emulated reference's; the bound must hold with room."""
    import cmvn_cases
    scale, zp = np.float32(0.046360891312360764), -11
    inv_scale = np.float32(1.0 / float(scale))
    src = pad_rows()
    mats = cmvn_cases.adversarial_matrices(scale)  # the GPU test feeds the same 72 matrices to the device code
    assert len(mats) == 72
    worst, certified, total = 0.0, 0, 0
    for mi, F32 in enumerate(mats):
        f_ref = reference_cmvn_f32(F32)
        with np.errstate(all="ignore"):
            t_ref = (f_ref / scale).astype(np.float32)
            k_ref = np.where(t_ref >= 0, np.floor(t_ref.astype(np.float64) + 0.5), np.ceil(t_ref.astype(np.float64) - 0.5))
        G = F32.astype(np.float64)[src]
        PS = np.concatenate([np.zeros((1, 13)), np.cumsum(G, axis=0)])
        PQ = np.concatenate([np.zeros((1, 13)), np.cumsum(G * G, axis=0)])
        S, Q = PS[101:150] - PS[0:49], PQ[101:150] - PQ[0:49]
        ok, k, tc, B = certify(S, Q, Q, F32, inv_scale)
        assert not np.any(ok & (k_ref != k)), f"matrix {mi}: level 1 certified a wrong rounding"
        with np.errstate(all="ignore"):
            r = (np.abs(t_ref.astype(np.float64) - tc.astype(np.float64)) / B.astype(np.float64))[ok]
        if r.size:
            worst = max(worst, float(r.max()))
        certified += int(ok.sum())
        total += ok.size
        # level 2 with the reference's own float mean
        acc = np.zeros((49, 13), np.float32)
        for w in range(101):
            acc = (acc + G[np.arange(49) + w].astype(np.float32)).astype(np.float32)
        mean = (acc / np.float32(101.0)).astype(np.float32)
        M = S / 101.0
        dm = mean.astype(np.float64) - M
        V2 = (Q - S * M) + 101.0 * dm * dm
        S2 = mean.astype(np.float64) * 101.0
        ok2, k2, tc2, B2 = certify(S2, V2 + S2 * (S2 / 101.0), Q, F32, inv_scale, em_term=False)
        assert not np.any(ok2 & (k_ref != k2)), f"matrix {mi}: level 2 certified a wrong rounding"
        with np.errstate(all="ignore"):
            r2 = (np.abs(t_ref.astype(np.float64) - tc2.astype(np.float64)) / B2.astype(np.float64))[ok2]
        if r2.size:
            worst = max(worst, float(r2.max()))
    assert worst < 0.6, f"reference within {worst:.2f} B of the bound"
    assert certified > 0.3 * total  # the guard conditions must not simply refuse everything


def test_cmvn_stage_of_the_port_is_the_reference_on_matrices_no_clip_produces():
    """pins PortOracle.cmvn_quantise -- the CPU side of the GPU test of the device shortcut -- to the UNMODIFIED reference's own
    processing::cmvnw (oracle/_ref) on the adversarial, special-value and near-boundary matrices themselves"""
    import cmvn_cases
    from oracle_lib import RefOracle, have_ref
    if not have_ref("l476"):
        pytest.skip("oracle/_ref not built")
    scale = np.float32(0.046360891312360764)
    mats = np.stack(cmvn_cases.adversarial_matrices(scale) + cmvn_cases.special_value_matrices() +
                    list(cmvn_cases.near_boundary_matrices(64, scale, seed=5)[0]))
    with np.errstate(all="ignore"):
        _, f_port = PortOracle("l476").cmvn_quantise(mats, want_features=True)
        f_ref = RefOracle("l476").cmvnw(mats)
    assert np.array_equal(f_port.view(np.uint32), f_ref.view(np.uint32)) or np.all((f_port == f_ref) | (np.isnan(f_port) & np.isnan(f_ref)))


def test_near_boundary_generator_lands_near_boundaries():
    import cmvn_cases
    scale = np.float32(0.046360891312360764)
    F, r = cmvn_cases.near_boundary_matrices(256, scale, seed=11)
    _, f = PortOracle("l476").cmvn_quantise(F, want_features=True)
    t = (f.reshape(-1, 49, 13) / scale)[np.arange(256)[:, None], r, np.arange(13)[None, :]].astype(np.float64)
    dist = np.abs(t - np.floor(t) - 0.5)
    assert np.median(dist) < 1e-3 and (dist < 1e-5).mean() > 0.1  # targeted chains sit on the boundary at rounding-error scale
