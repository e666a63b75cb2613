"""Synthetic DIA-MS slice pools (no reference counterpart — the reference trains on real sqMass-derived
slices).  Shapes and value ranges follow the reference's notebooks (MS2 `(N, 34, 40000)`
integer-valued and mostly zero, MS1 `(N, 34)` chromatogram with intensities ~8e4-2e5) and the recipe in
SURVEY.md §8d.  numpy only, seeded, so tests, bench and the golden generator build identical pools."""
import numpy as np


def synth_pool(n, rt=34, mz=40000, seed=1234, density=0.02):
    """Returns (ms2 int32 (n, rt, mz), ms1 int32 (n, rt)).

    MS2: Bernoulli(density) support per m/z column x Gaussian elution profile over RT
    (centre U[0, rt), sigma U[2, 6]) x Exp(scale 2000), floored to int, >= 0, max > min guaranteed.
    MS1: Gaussian chromatogram x U[5e4, 7e5] + U[0, 5e3] baseline.
    """
    rng = np.random.default_rng(seed)
    ms2 = np.zeros((n, rt, mz), dtype=np.int32)
    ms1 = np.zeros((n, rt), dtype=np.int32)
    r = np.arange(rt, dtype=np.float64)
    for i in range(n):
        support = np.nonzero(rng.random(mz) < density)[0]
        if support.size == 0:
            support = np.array([int(rng.integers(0, mz))])
        centre = rng.uniform(0, rt, size=support.size)
        sigma = rng.uniform(2.0, 6.0, size=support.size)
        height = rng.exponential(2000.0, size=support.size) + 1.0
        prof = np.exp(-0.5 * ((r[:, None] - centre[None, :]) / sigma[None, :]) ** 2) * height[None, :]
        ms2[i][:, support] = np.floor(prof).astype(np.int32)
        if ms2[i].max() == ms2[i].min():
            ms2[i, rt // 2, support[0]] = 1000
        c1 = rng.uniform(0, rt)
        s1 = rng.uniform(2.0, 6.0)
        chrom = np.exp(-0.5 * ((r - c1) / s1) ** 2) * rng.uniform(5e4, 7e5) + rng.uniform(0, 5e3, size=rt)
        ms1[i] = np.floor(chrom).astype(np.int32)
    return ms2, ms1
