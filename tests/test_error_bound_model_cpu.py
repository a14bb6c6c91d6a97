"""The completeness argument of the exact re-rank (wisecondorx_b200/csrc/rerank.cu: eps_from / approx_eps / cand_eps
and the per-candidate refinement), restated in NumPy and checked on emulated sweep arithmetic: operands rounded to f16
after centring and power-of-two scaling (newref_prep.cu), fp32 accumulation, fp32 norms and list values.

Checked per target row: (1) every selected candidate's approximate value is within its own bound eps_j of the exact
one; (2) the exact top-k (reference order: distance, then position) lies inside {v <= v_(k) + 2 eps} and survives the
refinement v_j - eps_j <= max over the k smallest v of (v_j + eps_j).  This is a model check of the bound, not of the
CUDA code -- the GPU tests compare the kernels' output with the oracle bit for bit."""
import numpy as np
import pytest

from wisecondorx_b200 import synth

K_PAD_UNIT = 64       # f16 K block
ABS_ERR = 2.0 ** -14  # api.cu: WCX_F16_ABS_ERR
U_WORST = 4.8852e-4


def eps_from(an, nb, e, s_d, k_pad):
    na = np.sqrt(an)
    gamma = (k_pad + 64.0) * 1.1920929e-7
    return 1.5 * (2.0 * s_d * e + e * e + 2.0 * gamma * na * nb + 4.77e-7 * (an + nb * nb + 2.0 * na * nb)) + 1e-300


@pytest.mark.parametrize("seed,noise,quantise", [(1, 0.05, False), (2, 0.2, False), (3, 0.01, True)])
def test_exact_topk_inside_refined_selection(seed, noise, quantise):
    per = [700, 600, 500, 400, 300, 200] + [60] * 16
    x, per, cum = synth.make_corrected_matrix(per, 96, seed=seed, noise=noise)
    if quantise:
        x = np.round(x * 64) / 64  # ties and exactly representable values
    n, s = x.shape
    k = 60
    k_pad = -(-s // K_PAD_UNIT) * K_PAD_UNIT
    mean = x.mean(0)
    q = np.frexp(np.abs(x).max() + np.abs(mean).max())[1]
    sc = 2.0 ** (14 - q)
    xs = (x - mean) * sc
    xh16 = xs.astype(np.float16)
    xh = xh16.astype(np.float64)
    norm = (xh ** 2).sum(1).astype(np.float32).astype(np.float64)
    resn = np.sqrt(((xs - xh) ** 2).sum(1)) * (1 + 1e-9)
    tau = 2.0 * np.sqrt(k_pad) * ABS_ERR
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = np.where(norm > 0, np.maximum(resn - tau, 0.0) / np.sqrt(norm), 0.0)
    rho_max = min(float(rho.max()) * (1 + 1e-6), U_WORST)
    xh32 = xh16.astype(np.float32)
    chr_of = np.searchsorted(cum, np.arange(n), side="right")
    rng = np.random.default_rng(seed)
    for r in rng.choice(n, 40, replace=False):
        c = chr_of[r]
        cs, ce = (0 if c == 0 else cum[c - 1]), cum[c]
        acc = (xh32 @ xh32[r]).astype(np.float32)
        v = (norm.astype(np.float32) - np.float32(2.0) * acc).astype(np.float64)
        v[cs:ce] = np.inf
        d = ((x - x[r]) ** 2).sum(1) * sc * sc  # exact distances in the units of the list values
        d[cs:ce] = np.inf
        an, ra = norm[r], resn[r]
        na = np.sqrt(an)
        vk = np.sort(v)[k - 1]
        big_d = max(vk + an, 0.0)
        dq = big_d * 1.02 + 1e-300
        for _ in range(4):  # fixed point: the bound must hold for distances up to D + eps
            s_d = np.sqrt(dq)
            nb = na + s_d
            eps_u = eps_from(an, nb, ra + rho_max * nb + 2.0 * tau, s_d, k_pad)
            if big_d + eps_u <= dq:
                break
            dq = (big_d + eps_u) * 1.05
        assert big_d + eps_u <= dq
        sel = np.flatnonzero(v <= vk + 2.0 * eps_u)
        vj = v[sel]
        s_dj = np.sqrt(np.maximum(vj + an, 0.0) * 1.02 + 1e-300)
        eps_j = eps_from(an, np.sqrt(norm[sel]), ra + resn[sel] + tau, s_dj, k_pad)
        assert (np.abs(vj - (d[sel] - an)) <= eps_j).all()
        assert (eps_j <= eps_u * (1 + 1e-12)).all()
        upper = (vj + eps_j)[vj <= vk].max()
        keep = set(sel[vj - eps_j <= upper].tolist())
        order = np.lexsort((np.arange(n), d))  # (distance, position): the reference's insertion order
        truth = [j for j in order[:k + (ce - cs)] if not (cs <= j < ce)][:k]
        assert set(truth) <= keep
