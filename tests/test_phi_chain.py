"""The clamped-affine accumulator (csrc/pt_math.cuh phi_compose / phi_apply) against the recursion it
replaces: castRay level d returns  A_d + clamp(0, 5, (L_{d+1} * f_d))  with clamp = max(lo, min(hi, v)) (NaN -> hi,
src/global.hpp:16-18); the outermost level is not clamped.  Scripted (A, f, terminal) chains incl. negative, zero, huge,
infinite and NaN factors.  CPU-only (the same header compiled for the host)."""
import ctypes as C

import numpy as np

import support as S


def clamp(lo, hi, v):
    m = np.where(v < hi, v, hi)  # std::min(hi, v): NaN -> hi
    return np.where(lo < m, m, lo)


def recursion(levels, terminal):
    """Inner-to-outer evaluation, float32 throughout, the way the reference unwinds."""
    x = terminal.astype(np.float32)
    depth = levels.shape[1]
    with np.errstate(all="ignore"):
        for d in range(depth - 1, -1, -1):
            A, f = levels[:, d, 0], levels[:, d, 1]
            x = (A + clamp(np.float32(0), np.float32(5), (x * f).astype(np.float32))).astype(np.float32)
    return x


def run(levels, terminal):
    n, depth = levels.shape[:2]
    out = np.zeros(n, np.float32)
    lv = np.ascontiguousarray(levels, np.float32)
    tm = np.ascontiguousarray(terminal, np.float32)
    S.hc_lib().hc_phi_chain(S.fp(lv), S.fp(tm), C.c_long(n), depth, S.fp(out))
    return out


def test_ordinary_chains_match_to_rounding():
    rng = np.random.RandomState(1)
    for depth in (1, 2, 3, 6, 12):
        n = 20000
        A = np.clip(rng.exponential(2.0, (n, depth)), 0, 15).astype(np.float32)       # A_d = clamp(0, 15, l_dir)
        f = (rng.exponential(0.8, (n, depth)) * (rng.rand(n, depth) < 0.9)).astype(np.float32)  # some zero factors
        term = rng.exponential(1.0, n).astype(np.float32) * (rng.rand(n) < 0.8)
        lv = np.stack([A, f], -1)
        got, want = run(lv, term), recursion(lv, term)
        assert np.allclose(got, want, rtol=2e-5, atol=1e-5), depth


def test_negative_and_raw_terminals():
    """RR-terminated vertices return the raw l_dir (may be negative); pdf can be negative (signed N.h) -> negative factors."""
    rng = np.random.RandomState(2)
    n, depth = 20000, 4
    A = np.clip(rng.exponential(2.0, (n, depth)), 0, 15).astype(np.float32)
    f = rng.normal(0, 1.5, (n, depth)).astype(np.float32)
    term = rng.normal(0, 3, n).astype(np.float32)
    lv = np.stack([A, f], -1)
    got, want = run(lv, term), recursion(lv, term)
    assert np.allclose(got, want, rtol=5e-5, atol=2e-5)


def test_nan_factors_and_nan_terminals():
    """NaN factor (0/0: eval and pdf both zero) -> the level is A + 5; NaN terminal under non-negative factors -> upper bound."""
    rng = np.random.RandomState(3)
    n, depth = 5000, 3
    A = np.clip(rng.exponential(2.0, (n, depth)), 0, 15).astype(np.float32)
    f = rng.exponential(0.8, (n, depth)).astype(np.float32)
    term = rng.exponential(1.0, n).astype(np.float32)
    f[rng.rand(n, depth) < 0.15] = np.nan
    lv = np.stack([A, f], -1)
    got, want = run(lv, term), recursion(lv, term)
    assert np.allclose(got, want, rtol=2e-5, atol=1e-5)
    term2 = term.copy()
    term2[::3] = np.nan
    f2 = np.abs(np.nan_to_num(f, nan=1.0)).astype(np.float32)
    lv2 = np.stack([A, f2], -1)
    assert np.allclose(run(lv2, term2), recursion(lv2, term2), rtol=2e-5, atol=1e-5)


def test_infinite_factors_with_non_negative_radiance():
    """f = +inf (pdf == 0 with eval > 0): the reference gives A + 5 for L > 0 and for L == 0 (0 * inf = NaN -> 5); the map
    treats the level as the constant A + 5.  (Only L < 0 — a negative raw l_dir further down — would differ; documented.)"""
    rng = np.random.RandomState(4)
    n, depth = 5000, 3
    A = np.clip(rng.exponential(2.0, (n, depth)), 0, 15).astype(np.float32)
    f = rng.exponential(0.8, (n, depth)).astype(np.float32)
    f[rng.rand(n, depth) < 0.2] = np.inf
    term = (rng.exponential(1.0, n) * (rng.rand(n) < 0.7)).astype(np.float32)  # >= 0, some exactly 0
    lv = np.stack([A, f], -1)
    assert np.allclose(run(lv, term), recursion(lv, term), rtol=2e-5, atol=1e-5)


def test_depth_zero_is_the_identity_including_nan():
    term = np.array([0.0, -1.5, 7.25, np.nan, np.inf], np.float32)
    got = run(np.zeros((5, 0, 2), np.float32), term)
    assert np.array_equal(np.isnan(got), np.isnan(term)) and np.array_equal(got[~np.isnan(term)], term[~np.isnan(term)])


def test_settled_paths_end_early_with_the_same_bits():
    """shade_kernel ends a path at a level d >= 1 when the map composed so far takes the same value at A_d and at A_d + 5 (the
    level's value lies in between whatever follows).  The result must be the bits the full chain gives, for ordinary, negative,
    zero, infinite and NaN factors and terminals, and the shortcut must actually fire on chains with brightly lit vertices."""
    rng = np.random.RandomState(9)
    n, depth = 40000, 6
    A = np.clip(rng.exponential(3.0, (n, depth)), 0, 15).astype(np.float32)
    A[rng.rand(n, depth) < 0.3] = 0
    f = (rng.exponential(1.2, (n, depth)) * np.where(rng.rand(n, depth) < 0.1, -1, 1)).astype(np.float32)
    f[rng.rand(n, depth) < 0.05] = 0
    f[rng.rand(n, depth) < 0.03] = np.inf
    f[rng.rand(n, depth) < 0.03] = np.nan
    term = rng.normal(1.0, 3.0, n).astype(np.float32)
    term[rng.rand(n) < 0.05] = np.nan
    term[rng.rand(n) < 0.02] = np.inf
    lv = np.ascontiguousarray(np.stack([A, f], -1), np.float32)
    full = run(lv, term)
    out = np.zeros(n, np.float32)
    where = np.zeros(n, np.int32)
    S.hc_lib().hc_phi_chain_settled(S.fp(lv), S.fp(np.ascontiguousarray(term)), C.c_long(n), depth, S.fp(out), S.ip(where))
    same = (out.view(np.uint32) == full.view(np.uint32)) | (np.isnan(out) & np.isnan(full))
    assert same.all(), f"{(~same).sum()} chains differ"
    assert (where >= 1).mean() > 0.2, (where >= 1).mean()
