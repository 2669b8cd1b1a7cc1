"""Executable specification (CPU, no GPU) of the arithmetic behind the exact parallel ordered fold
(starr_b200/csrc/st_ordered_fold.cuh: scan_ready / decompose / pmap_then / value_of / walk_updates).

The reference folds a node's updates one IEEE add at a time (_sumtree_array.pyx:46-48).  The CUDA path
replaces runs of those adds by integer maps m -> m + a[m & 1] that compose associatively, and falls
back to a real add exactly where the running value would leave its binade.  This file restates that
arithmetic with Python integers and checks it against NumPy's sequential float adds -- every
intermediate value, bit for bit -- so the claim "nothing is approximate" is checked on the CPU as well
(the GPU parity tests check the kernels themselves)."""
import numpy as np
import pytest

FMT = {np.float32: (np.uint32, 23, 0xFF, 32), np.float64: (np.uint64, 52, 0x7FF, 64)}


def bits(x, dt):
    return int(np.array([x], dt).view(FMT[dt][0])[0])


def from_bits(b, dt):
    return np.array([b], FMT[dt][0]).view(dt)[0]


def scan_ready(v, dt):
    """(ok, eT, m): positive normal v = m * 2^(eT - bias - MB) with m in [2^MB, 2^(MB+1))."""
    _, mb, emask, nbits = FMT[dt]
    b = bits(v, dt)
    ef = (b >> mb) & emask
    m = (b & ((1 << mb) - 1)) | (1 << mb)
    return (b >> (nbits - 1)) == 0 and 1 <= ef < emask, ef, m


def value_of(et, m, dt):
    mb = FMT[dt][1]
    return from_bits((et << mb) + (m - (1 << mb)), dt)


def decompose(d, et, dt):
    """(a0, a1, flq, big) of one update seen from a node with exponent field et; flq = floor(4 d / ulp)."""
    _, mb, emask, nbits = FMT[dt]
    b = bits(d, dt)
    neg = (b >> (nbits - 1)) != 0
    ef = (b >> mb) & emask
    mant = b & ((1 << mb) - 1)
    if ef == emask:
        return 0, 0, 0, True
    md = (mant | (1 << mb)) if ef else mant
    if md == 0:
        return 0, 0, 0, False
    shift = (ef if ef else 1) - et
    if shift >= 0:
        if shift >= 2:
            return 0, 0, 0, True
        q = md << shift
        g0 = g1 = q
        q4, frac4 = q << 2, False
    else:
        sh = -shift
        if sh > mb + 2:
            g0 = g1 = 0
            q4, frac4 = 0, True
        else:
            q = md >> sh
            rem = md & ((1 << sh) - 1)
            half = 1 << (sh - 1)
            if rem > half:
                g0 = g1 = q + 1
            elif rem == half:                      # exact tie: round-half-even depends on the parity of m
                g0, g1 = q + (q & 1), q + ((q + 1) & 1)
            else:
                g0 = g1 = q
            if sh >= 2:
                q4, frac4 = md >> (sh - 2), (md & ((1 << (sh - 2)) - 1)) != 0
            else:
                q4, frac4 = md << 1, False
    if not neg:
        return g0, g1, q4, False
    return -g0, -g1, -q4 - (1 if frac4 else 0), False


def covered(m, flq, dt):
    """walk_updates' test: the identity holds while the exact sum stays in [2^MB - 1/4, 2^(MB+1)) ulps -- the binade
    itself plus the quarter below it, where the finer grid of the lower binade rounds to 2^MB as well."""
    mb = FMT[dt][1]
    lo, hi = 1 << mb, 1 << (mb + 1)
    return lo <= m < hi and 4 * (m - lo) + min(flq, 0) >= -1 and m + (flq >> 2) < hi


def then(f, g):
    """apply f first, then g (pmap_then)."""
    return (f[0] + (g[1] if f[0] & 1 else g[0]), f[1] + (g[1] if (f[1] + 1) & 1 else g[0]))


def fold_with_maps(t, ds, dt):
    """The scan-with-restart of ordered_fold_*: maps while the exact partial sums stay inside the
    binade, one real add where they would not (or where the value is not a positive normal)."""
    mb = FMT[dt][1]
    lo, hi = 1 << mb, 1 << (mb + 1)
    cur = dt(t)
    pos, real_adds = 0, 0
    while pos < len(ds):
        ok, et, m0 = scan_ready(cur, dt)
        if not ok:
            cur = dt(cur + ds[pos])
            pos += 1
            real_adds += 1
            continue
        # prefix composition = what the warp / block scan computes for every position at once
        pre = (0, 0)
        viol = None
        for k in range(pos, len(ds)):
            a0, a1, flq, big = decompose(ds[k], et, dt)
            m = m0 + (pre[1] if m0 & 1 else pre[0])            # exact m in front of update k
            if big or not covered(m, flq, dt):                 # walk_updates' test
                viol = (k, m)
                break
            pre = then(pre, (a0, a1))
        if viol is None:
            return value_of(et, m0 + (pre[1] if m0 & 1 else pre[0]), dt), real_adds
        k, m = viol
        cur = dt(value_of(et, m, dt) + ds[k])                  # the violating update: one real IEEE add
        real_adds += 1
        pos = k + 1
    return cur, real_adds


def sequential(t, ds, dt):
    cur = dt(t)
    for d in ds:
        cur = dt(cur + d)
    return cur


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
def test_single_update_map_equals_ieee_add(dt):
    """fl(T + d) == u * (m + a[m & 1]) whenever the exact sum stays in T's binade -- all magnitudes of d
    relative to ulp(T), both signs, exact ties on both parities."""
    rng = np.random.default_rng(1)
    mb = FMT[dt][1]
    checked = 0
    for _ in range(4000):
        t = dt(np.exp(rng.uniform(-20, 20)))
        ok, et, m = scan_ready(t, dt)
        assert ok
        ulp = float(value_of(et, m + 1, dt)) - float(t) if m + 1 < (1 << (mb + 1)) else float(t) - float(value_of(et, m - 1, dt))
        kind = rng.integers(0, 5)
        if kind == 4:                                  # bottom of the binade, updates of a fraction of an ulp
            t = value_of(et, (1 << mb) + int(rng.integers(0, 3)), dt)
            ok, et, m = scan_ready(t, dt)
            ulp = float(value_of(et, m + 1, dt)) - float(t)
            d = dt(-ulp * rng.choice([0.125, 0.25, 0.2500001, 0.375, 0.5, 0.75, 1.0, 1.25, 2.25, 1e-6]) * rng.choice([1, 1, 1, -1]))
        elif kind == 0:
            d = dt(rng.choice([-1, 1]) * ulp * rng.choice([0.5, 1.5, 2.5, 0.25, 0.75, 1.0, 3.0]))   # ties and near-ties
        elif kind == 1:
            d = dt(rng.choice([-1, 1]) * ulp * np.exp(rng.uniform(-30, 3)))
        elif kind == 2:
            d = dt(rng.choice([-1, 1]) * float(t) * rng.uniform(0, 0.4))
        else:
            d = dt(0.0)
        a0, a1, flq, big = decompose(d, et, dt)
        if big:
            continue
        if not covered(m, flq, dt):
            continue                                   # leaves the binade: the kernel uses a real add there
        got = value_of(et, m + (a1 if m & 1 else a0), dt)
        want = dt(t + d)
        assert bits(got, dt) == bits(want, dt), (t, d)
        checked += 1
    assert checked > 2000


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("kind", ["small", "ties", "hopping", "on_power_of_two", "from_zero", "mixed"])
def test_composed_maps_equal_sequential_fold(dt, kind):
    """Prefix composition + restart reproduces the sequential IEEE fold bit for bit, including runs that
    hop across binade boundaries, grow from zero, or consist of exact ties."""
    rng = np.random.default_rng(7)
    for rep in range(30):
        n = int(rng.integers(1, 400))
        if kind == "small":
            t, ds = 1.5 * 2.0 ** rng.integers(0, 14), rng.normal(0, 1e-3, n)   # mid-binade: never leaves it
        elif kind == "ties":
            t = 2.0 ** rng.integers(10, 20) + rng.integers(0, 64)
            ulp = float(np.spacing(dt(t)))
            ds = rng.choice([-1.5, -0.5, 0.5, 1.5, 2.5], n) * ulp
        elif kind == "hopping":                       # sits on a power of two and walks across it
            t = 2.0 ** rng.integers(5, 15)
            ds = rng.normal(0, float(np.spacing(dt(t))) * 3, n)
        elif kind == "on_power_of_two":               # every update is absorbed: the node never leaves 2^k
            t = 2.0 ** rng.integers(5, 15)
            ds = rng.uniform(-0.249, 0.49, n) * float(np.spacing(dt(t)))
        elif kind == "from_zero":
            t, ds = 0.0, np.exp(rng.uniform(-30, 5, n))
        else:
            t = np.exp(rng.uniform(-5, 5))
            ds = np.exp(rng.uniform(-25, 6, n)) * rng.choice([-1, 1], n) * (rng.random(n) < 0.9)
            ds[rng.integers(0, n)] = -t * 0.999      # near-cancellation
        ds = ds.astype(dt)
        got, real_adds = fold_with_maps(t, ds, dt)
        want = sequential(t, ds, dt)
        assert bits(got, dt) == bits(want, dt), (kind, rep, t)
        if kind in ("small", "on_power_of_two"):
            assert real_adds == 0                     # the whole run is one composition


def test_map_composition_is_associative():
    rng = np.random.default_rng(3)
    for _ in range(2000):
        f, g, h = [tuple(int(x) for x in rng.integers(-50, 50, 2)) for _ in range(3)]
        # parity maps of real updates satisfy |a0 - a1| <= 1; composition keeps that and is associative
        f, g, h = [(a, a + int(rng.integers(-1, 2))) for a, _ in (f, g, h)]
        assert then(then(f, g), h) == then(f, then(g, h))
        for m in (1000, 1001):
            step = lambda mm, p: mm + (p[1] if mm & 1 else p[0])
            assert step(step(step(m, f), g), h) == step(m, then(then(f, g), h))


def chunk_record(ds, et, dt):
    """chunk_summary_cta: parity map of the whole chunk plus, for either start parity, the extremes of
    4 (m_k - m_in) + min(floor(4 x_k), 0)   (QUARTER units)   and   (m_k - m_in) + max(floor(x_k), 0)."""
    rec = {"ok": True}
    maps = []
    for d in ds:
        a0, a1, flq, big = decompose(d, et, dt)
        rec["ok"] &= not big
        maps.append((a0, a1, flq))
    tot = (0, 0)
    for a0, a1, _ in maps:
        tot = then(tot, (a0, a1))
    rec["a"] = tot
    for parity in (0, 1):
        off, mn, mx = 0, 0, 0
        for a0, a1, flq in maps:
            mn = min(mn, 4 * off + min(flq, 0))
            mx = max(mx, off + (flq >> 2) if flq > 0 else off)
            off += a1 if (off + parity) & 1 else a0
        rec[parity] = (mn, mx)
    return rec


def absorbed_chunk_record(ds, et, dt):
    """The early-out of chunk_summary_cta: every element's map is the identity -> no scan, two reductions."""
    flqs = []
    ok = True
    for d in ds:
        a0, a1, flq, big = decompose(d, et, dt)
        assert a0 == 0 and a1 == 0
        ok &= not big
        flqs.append(flq)
    mn = min([0] + [f for f in flqs if f < 0])
    mx = max([0] + [f >> 2 for f in flqs if f > 0])
    return {"ok": ok, "a": (0, 0), 0: (mn, mx), 1: (mn, mx)}


@pytest.mark.parametrize("dt", [np.float32, np.float64], ids=["f32", "f64"])
def test_chunk_records_cover_exactly_what_the_element_walk_covers(dt):
    """A chunk record accepts an incoming m (4 (m - LO) + min >= -1 and m + max < HI) exactly when walking its
    elements one by one never meets an update the identity does not cover; then m_out = m + a[parity]; and the
    absorbed-chunk shortcut produces the same record as the general path."""
    rng = np.random.default_rng(11)
    mb = FMT[dt][1]
    lo, hi = 1 << mb, 1 << (mb + 1)
    for rep in range(200):
        t = dt(2.0 ** rng.integers(3, 12) * rng.choice([1.0, 1.0 + 2.0 ** -mb, 1.5, 2.0 - 2.0 ** -mb]))   # on / next to a boundary, mid-binade
        ok, et, m0 = scan_ready(t, dt)
        ulp = float(np.spacing(t))
        scale = rng.choice([0.2, 0.6, 3.0])
        ds = (rng.uniform(-1, 1, int(rng.integers(1, 64))) * ulp * scale).astype(dt)
        rec = chunk_record(ds, et, dt)
        if all(decompose(d, et, dt)[:2] == (0, 0) for d in ds):
            assert absorbed_chunk_record(ds, et, dt) == rec
        for m in (m0, lo, lo + 1, hi - 1, hi - 2, (lo + hi) // 2):
            mn, mx = rec[m & 1]
            accept = rec["ok"] and 4 * (m - lo) + mn >= -1 and m + mx < hi
            mm, walk_ok = m, True
            for d in ds:
                a0, a1, flq, big = decompose(d, et, dt)
                if big or not covered(mm, flq, dt):
                    walk_ok = False
                    break
                mm += a1 if mm & 1 else a0
            assert accept == walk_ok, (rep, m - lo)
            if accept:
                assert mm == m + (rec["a"][1] if m & 1 else rec["a"][0])
                want = sequential(value_of(et, m, dt), ds, dt)
                assert bits(value_of(et, mm, dt), dt) == bits(want, dt)
