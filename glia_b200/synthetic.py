"""Synthetic atlas-shaped inputs for benchmarks and tests.

Pure input generation on the host (NumPy/SciPy); nothing here is on the solver path.
Everything is a deterministic function of (n, seed), so that the GPU arm, the CPU reference
arm and the tests see bit-identical inputs.

  * three smooth random fields (white noise restricted to |w| <= 8) -> soft-max into WM / GM /
    CSF probability maps inside an ellipsoidal "brain" (semi-axes 0.38/0.45/0.36 of the box),
    background outside, an ellipsoidal ventricle (VT) in the middle;
  * every map smoothed with a periodic Gaussian of sigma = 2 pi / n and clipped to [0, 1];
  * brain filter as in the reference, (wm > 0.1 or gm > 0.1) and vt < 0.8
    (src/mat/MatProp.cpp:180-185);
  * initial condition: 1-3 Gaussians of sigma = 2 pi / 64 at seeded white-matter locations,
    activations p ~ U(0.5, 1), rescaled to max 1 (a TIL parametrisation, src/mat/Phi.cpp:264-374).
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft


def _axes(n):
    return [2.0 * np.pi * np.arange(m) / m for m in n]


def _lowpass_noise(n, rng, wmax=8):
    """Real periodic field whose spectrum lives on |w_d| <= wmax."""
    spec = np.zeros((n[0], n[1], n[2] // 2 + 1), np.complex128)
    m = wmax
    blk = rng.standard_normal((2 * m + 1, 2 * m + 1, m + 1)) + 1j * rng.standard_normal((2 * m + 1, 2 * m + 1, m + 1))
    w = np.arange(-m, m + 1)
    amp = 1.0 / (1.0 + (w[:, None, None] ** 2 + w[None, :, None] ** 2 + np.arange(m + 1)[None, None, :] ** 2))
    blk *= amp
    ix = np.arange(-m, m + 1) % n[0]
    iy = np.arange(-m, m + 1) % n[1]
    spec[np.ix_(ix, iy, np.arange(m + 1))] = blk
    f = sfft.irfftn(spec, s=tuple(n), workers=-1)
    return (f - f.mean()) / f.std()


def _gauss_smooth(f, sigma):
    n = f.shape
    w = [np.fft.fftfreq(m, 1.0 / m) for m in n]
    fh = sfft.rfftn(f, workers=-1)
    g = (np.exp(-0.5 * sigma ** 2 * w[0] ** 2)[:, None, None] * np.exp(-0.5 * sigma ** 2 * w[1] ** 2)[None, :, None]
         * np.exp(-0.5 * sigma ** 2 * w[2][: n[2] // 2 + 1] ** 2)[None, None, :])
    return sfft.irfftn(fh * g, s=n, workers=-1)


def _upsample2(a):
    """periodic linear interpolation to twice the resolution along every axis"""
    for ax in range(a.ndim):
        nxt = np.roll(a, -1, axis=ax)
        out = np.empty(a.shape[:ax] + (2 * a.shape[ax],) + a.shape[ax + 1:], a.dtype)
        idx = [slice(None)] * a.ndim
        idx[ax] = slice(0, None, 2)
        out[tuple(idx)] = a
        idx[ax] = slice(1, None, 2)
        out[tuple(idx)] = 0.5 * (a + nxt)
        a = out
    return a


def make_atlas(n, seed=0, dtype=np.float32):
    """-> dict(wm, gm, csf, vt, filter) of C-ordered [n0][n1][n2] arrays in `dtype`.
    Grids above 256^3 are the 256^3 atlas interpolated (periodic, linear) to the finer grid: the
    same anatomy at every resolution, and no minute-long host FFTs before a multi-GPU run."""
    n = (n, n, n) if np.isscalar(n) else tuple(int(v) for v in n)
    if min(n) > 256 and all(v % 2 == 0 for v in n):
        coarse = make_atlas(tuple(v // 2 for v in n), seed, np.float32)
        out = {k: np.ascontiguousarray(_upsample2(coarse[k]).astype(dtype)) for k in ("wm", "gm", "csf", "vt")}
        out["filter"] = (((out["wm"] > 0.1) | (out["gm"] > 0.1)) & (out["vt"] < 0.8)).astype(dtype)
        return out
    rng = np.random.default_rng(seed)
    ax = _axes(n)
    u = [(a - np.pi) / (2 * np.pi) for a in ax]  # box coordinates in [-0.5, 0.5)
    r_brain = (u[0][:, None, None] / 0.38) ** 2 + (u[1][None, :, None] / 0.45) ** 2 + (u[2][None, None, :] / 0.36) ** 2
    r_vt = (u[0][:, None, None] / 0.06) ** 2 + ((u[1][None, :, None] - 0.02) / 0.10) ** 2 + (u[2][None, None, :] / 0.05) ** 2
    brain = (r_brain < 1.0).astype(np.float64)
    vt = (r_vt < 1.0).astype(np.float64)
    logits = np.stack([1.5 * _lowpass_noise(n, rng) + b for b in (0.6, 0.0, -0.8)])
    logits -= logits.max(axis=0, keepdims=True)
    e = np.exp(logits)
    prob = e / e.sum(axis=0, keepdims=True)
    sigma = 2.0 * np.pi / n[0]
    maps = {}
    for name, p in zip(("wm", "gm", "csf"), prob):
        maps[name] = np.clip(_gauss_smooth(p * brain * (1.0 - vt), sigma), 0.0, 1.0)
    maps["vt"] = np.clip(_gauss_smooth(vt, sigma), 0.0, 1.0)
    tot = maps["wm"] + maps["gm"] + maps["csf"] + maps["vt"]
    scale = np.where(tot > 1.0, 1.0 / np.maximum(tot, 1e-30), 1.0)
    out = {k: np.ascontiguousarray((v * scale).astype(dtype)) for k, v in maps.items()}
    out["filter"] = (((out["wm"] > 0.1) | (out["gm"] > 0.1)) & (out["vt"] < 0.8)).astype(dtype)
    return out


def make_initial_condition(atlas, seed=0, n_gauss=3, dtype=np.float32):
    """Gaussian TIL initial condition c(0), max 1, centred at seeded white-matter voxels."""
    wm = atlas["wm"]
    n = wm.shape
    rng = np.random.default_rng(seed + 1000)
    cand = np.argwhere(wm > 0.6)
    if len(cand) == 0:
        cand = np.argwhere(wm >= wm.max() * 0.9)
    first = cand[rng.integers(len(cand))]
    # keep the other centres within a few sigma of the first so the tumour is one focus
    sig = 2.0 * np.pi / 64
    ax = _axes(n)
    c0 = np.zeros(n)
    for j in range(n_gauss):
        ctr = first if j == 0 else np.clip(first + rng.integers(-3, 4, 3) * np.maximum(np.array(n) // 64, 1),
                                           0, np.array(n) - 1)
        p = rng.uniform(0.5, 1.0)
        g = [np.exp(-(ax[d] - ax[d][ctr[d]]) ** 2 / (2 * sig * sig)) for d in range(3)]  # separable Gaussian
        c0 += p * (g[0][:, None, None] * g[1][None, :, None] * g[2][None, None, :])
    c0 *= atlas["filter"].astype(np.float64)
    c0 /= c0.max()
    return np.ascontiguousarray(c0.astype(dtype))
