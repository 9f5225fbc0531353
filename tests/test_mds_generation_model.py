"""CPU model of the exact-generation scheme of csrc/mds.cu, checked against the plain sequential algorithm (MDS_cuda.cu:91-211).

The CUDA kernel cuts the 16 383 dependent picks into generations: every "warp" (a fixed subset of the points) publishes its M
lowest (density, tie key) pairs and its (M+1)-th as a bound; theta = min of the bounds; the replay runs the sequential algorithm
on the pool alone and accepts a pick while (density, key) < theta.  The claim the kernel rests on -- an accepted pick IS the pick
of the sequential algorithm, including all tie cases, and at least one pick is accepted per generation -- is an algorithmic
invariant, so it is verified here in numpy with the SAME elementwise arithmetic on both sides (float32, one rounding per
addition), independent of any GPU.  A second model adds the round-2 plan (bounds refreshed while a generation runs) to show it
keeps the invariant."""
import numpy as np
import pytest


def _weights(xyz, p, t):
    d = xyz - xyz[p]
    d2 = (d[:, 2] * d[:, 2] + (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])).astype(np.float32)
    return np.exp(-(d2 / np.float32(t)).astype(np.float32)).astype(np.float32)


def _keys(n):
    bs = 1
    lg = 0
    while bs * 2 <= n and bs < 1024:
        bs *= 2
        lg += 1
    k = np.arange(n, dtype=np.int64)
    rev = np.zeros(n, dtype=np.int64)
    v = k & (bs - 1)
    for b in range(lg):
        rev |= ((v >> b) & 1) << (lg - 1 - b)
    return (rev << 21) | k            # tie key: bit-reversed thread slot of the reference's tournament, then the index


def sequential(xyz, m, t):
    n = len(xyz)
    fac = np.where(np.arange(n) < 8192, np.float32(1), np.float32(2))
    temp = np.zeros(n, dtype=np.float32)
    key = _keys(n)
    out = [0]
    temp[0] = np.float32(1e9)
    last = 0
    for _ in range(1, m):
        temp = (temp + _weights(xyz, last, t) * fac).astype(np.float32)
        temp[np.array(out)] = np.float32(1e9)
        order = np.lexsort((key, temp))
        last = int(order[0]) if temp[order[0]] < 1e9 else 0
        out.append(last)
        temp[last] = np.float32(1e9)
    return out


def generations(xyz, m, t, n_warps, M, refresh_every=0):
    """refresh_every > 0: the per-warp bounds are recomputed from the warps' CURRENT densities every that many accepted picks
    (a bound taken after pick j is valid for every later pick because densities only grow) -- the asynchronous-refresh plan."""
    n = len(xyz)
    fac = np.where(np.arange(n) < 8192, np.float32(1), np.float32(2))
    temp = np.zeros(n, dtype=np.float32)
    key = _keys(n)
    temp[0] = np.float32(1e9)
    temp = (temp + _weights(xyz, 0, t) * fac).astype(np.float32)
    temp[0] = np.float32(1e9)
    owner = np.arange(n) % n_warps                      # which "warp" holds a point (any fixed partition works)
    out = [0]
    n_gen = 0
    while len(out) < m:
        n_gen += 1
        pool, published = [], np.zeros(n, dtype=bool)
        bounds = []
        for w in range(n_warps):
            idx = np.nonzero(owner == w)[0]
            idx = idx[temp[idx] < 1e9]
            order = idx[np.lexsort((key[idx], temp[idx]))]
            pool.extend(order[:M].tolist())
            published[order[:M]] = True
            bounds.append((temp[order[M]], key[order[M]]) if len(order) > M else (np.float32(np.inf), 0))
        if not pool:
            out.extend([0] * (m - len(out)))
            break
        theta = min(bounds)
        ptemp = {p: temp[p] for p in pool}              # the replay works on its own copy of the pool's densities
        accepted = []
        while len(out) + len(accepted) < m:
            if accepted:
                w = _weights(xyz, accepted[-1], t)
                for p in ptemp:
                    ptemp[p] = np.float32(ptemp[p] + w[p] * fac[p])
            live = [p for p in ptemp if p not in accepted]
            if not live:
                break
            best = min(live, key=lambda p: (ptemp[p], key[p]))
            if refresh_every and accepted and len(accepted) % refresh_every == 0:
                cur = temp.copy()                       # what the workers hold after applying every accepted pick so far
                for a in accepted:
                    cur = (cur + _weights(xyz, a, t) * fac).astype(np.float32)
                    cur[a] = np.float32(1e9)
                un = np.nonzero(~published & (cur < 1e9))[0]
                if len(un):
                    o = un[np.lexsort((key[un], cur[un]))[0]]
                    theta = max(theta, (cur[o], key[o]))
                else:
                    theta = (np.float32(np.inf), 0)
            if not ((ptemp[best], key[best]) < theta):
                break
            accepted.append(best)
        assert accepted, "a generation must accept at least one pick (the global minimum is some warp's minimum)"
        for a in accepted:                              # the workers apply the picks in order
            temp = (temp + _weights(xyz, a, t) * fac).astype(np.float32)
            temp[np.array(out + [a])] = np.float32(1e9)
            out.append(a)
    return out[:m], n_gen


@pytest.mark.parametrize("n,m,mml,seed", [(300, 200, 0.05, 0), (700, 650, 0.02, 1), (520, 300, 0.3, 2), (9000, 120, 0.02, 3)])
def test_generations_reproduce_the_sequential_picks(n, m, mml, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.random((n, 3), dtype=np.float32)
    if seed == 1:
        xyz[600:] = 0                                   # identical padded points: exact density ties, decided by the tie key
    t = np.float32(5.0 * mml * mml)
    ref = sequential(xyz, m, t)
    for n_warps, M in ((8, 4), (32, 8), (5, 1)):
        got, n_gen = generations(xyz, m, t, n_warps, M)
        assert got == ref
        assert n_gen < m                                # generations really batch picks
    got, n_gen_refresh = generations(xyz, m, t, 8, 4, refresh_every=2)
    assert got == ref                                   # refreshed bounds keep the invariant ...
    assert n_gen_refresh <= generations(xyz, m, t, 8, 4)[1]   # ... and never shorten a generation


def test_oversampling_returns_index_zero():
    rng = np.random.default_rng(7)
    xyz = rng.random((40, 3), dtype=np.float32)
    t = np.float32(5.0 * 0.1 * 0.1)
    ref = sequential(xyz, 50, t)
    got, _ = generations(xyz, 50, t, 4, 2)
    assert got == ref and got[40:] == [0] * 10


@pytest.mark.parametrize("mml,seed", [(0.02, 0), (0.08, 1)])
def test_box_culling_criterion_never_skips_a_changing_point(mml, seed):
    """The slot test of the culling variant (csrc/mds.cu, CULL): a group of points may be skipped for a pick when
    2.02 * exp(-0.9999 * dist2(pick, box) / t) < min live density * 2^-25.  Then fl(density + w) == density must hold for every live
    point of the group -- checked here in float32 over a replay of the sampler, for groups of 32 in Z-order."""
    rng = np.random.default_rng(seed)
    n = 1536
    xyz = rng.random((n, 3), dtype=np.float32)
    t = np.float32(5.0 * mml * mml)
    fac = np.where(np.arange(n) < 700, np.float32(1), np.float32(2))     # both weights occur
    q = (xyz * 15.999).astype(np.int64)
    code = np.zeros(n, dtype=np.int64)
    for b in range(4):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    group = np.empty(n, dtype=np.int64)
    group[np.argsort(code, kind="stable")] = np.arange(n) // 32
    ng = n // 32
    lo = np.full((ng, 3), np.inf, dtype=np.float32)
    hi = np.full((ng, 3), -np.inf, dtype=np.float32)
    np.minimum.at(lo, group, xyz)
    np.maximum.at(hi, group, xyz)
    temp = np.zeros(n, dtype=np.float32)
    live = np.ones(n, dtype=bool)
    live[0] = False
    last, skipped = 0, 0
    for _ in range(400):
        w = (_weights(xyz, last, t) * fac).astype(np.float32)
        new = (temp + w).astype(np.float32)
        dd = np.maximum(np.maximum(lo - xyz[last], xyz[last] - hi), 0).astype(np.float32)
        dmin = ((dd * dd).sum(1) * np.float32(0.9999)).astype(np.float32)
        wmax = np.float32(2.02) * np.exp(-(dmin / t))
        tmin = np.full(ng, np.inf, dtype=np.float32)
        np.minimum.at(tmin, group[live], temp[live])
        skip = wmax < tmin * np.float32(2.0 ** -25)
        changed = (new != temp) & live
        assert not skip[group[changed]].any()
        skipped += int(skip[np.isfinite(tmin)].sum())
        temp = new
        cand = np.where(live, temp, np.float32(np.inf))
        last = int(np.lexsort((np.arange(n), cand))[0])
        live[last] = False
    if mml < 0.05:
        assert skipped > 0      # the criterion is not vacuous in the regime it is meant for
