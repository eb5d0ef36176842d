"""GPU parity: all-vs-all sketch comparison vs the oracle's literal compareSketches loop.
Shared-hash counts and denominators bit-exact; distance and p-value within 1e-12 relative."""
import numpy as np
import pytest

from util import sorted_sketch_panel

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.fixture(params=["rank32", "rank32_saturate", "rank32_noprune", "u64"])
def dctx(ctx, request):
    """The context with the tile kernel pinned: 32-bit codes with pruning of pairs that share no hash (default), the same with
    the marking walks bounded by the reference components (what large panels get), the same merging every pair, or the
    64-bit kernel."""
    ctx.set_dist_mode(force64=request.param == "u64", no_prune=request.param == "rank32_noprune", saturate=request.param == "rank32_saturate")
    yield ctx
    ctx.set_dist_mode()


def _oracle_matrix(oracle, ref, qry, s, k, kmer_space, max_d=1.0, max_p=1.0):
    rh, rs, rl = ref
    qh, qs, ql = qry
    out = []
    for q in range(len(qs)):
        row = []
        for r in range(len(rs)):
            row.append(oracle.compare(rh[r, :rs[r]], qh[q, :qs[q]], int(rl[r]), int(ql[q]), s, k, kmer_space, max_d, max_p))
        out.append(row)
    return out


def _compare(got, passed, want):
    for q, row in enumerate(want):
        for r, w in enumerate(row):
            g = got[q, r]
            assert bool(passed[q, r]) == w["passed"], (q, r)
            assert int(g["numer"]) == w["numer"] and int(g["denom"]) == w["denom"], (q, r, g, w)
            assert g["distance"] == pytest.approx(w["distance"], rel=RTOL, abs=0), (q, r)
            if w["passed"] or w["pvalue"] != 0:
                if w["pvalue"] == 0:
                    assert g["pvalue"] == 0
                else:
                    assert g["pvalue"] == pytest.approx(w["pvalue"], rel=RTOL, abs=1e-300), (q, r, g, w)


@pytest.mark.parametrize("s,n_ref,n_qry", [(1000, 37, 21), (1000, 64, 16), (100, 33, 50), (2500, 20, 9)])
def test_dist_sorted_panels(dctx, oracle, s, n_ref, n_qry):
    rng = np.random.default_rng(s + n_ref)
    rh, rs = sorted_sketch_panel(rng, n_ref, s)
    qh, qs = sorted_sketch_panel(rng, n_qry, s)
    qh[:3] = rh[:3]; qs[:3] = rs[:3]          # identical sketches: distance 0, 1000/1000
    rl = rng.integers(1000, 6_000_000, size=n_ref).astype(np.uint64)
    ql = rng.integers(1000, 6_000_000, size=n_qry).astype(np.uint64)
    got, passed = dctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21)
    _compare(got, passed, _oracle_matrix(oracle, (rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21))


def test_dist_thresholds_and_sketch_size_mismatch(dctx, oracle):
    rng = np.random.default_rng(77)
    rh, rs = sorted_sketch_panel(rng, 40, 1000, shared=0.8)
    qh, qs = sorted_sketch_panel(rng, 24, 1000, shared=0.8)
    rl = np.full(40, 4_600_000, dtype=np.uint64)
    ql = np.full(24, 5_100_000, dtype=np.uint64)
    # compare at s=400 although the lists hold up to 1000 hashes (CommandDistance.cpp:342-344)
    got, passed = dctx.dist_tile((rh, rs, rl), (qh, qs, ql), 400, 21, 4.0 ** 21, max_distance=0.06, max_pvalue=1e-3)
    _compare(got, passed, _oracle_matrix(oracle, (rh, rs, rl), (qh, qs, ql), 400, 21, 4.0 ** 21, 0.06, 1e-3))


def test_dist_heterogeneous_densities_multi_phase(dctx, oracle):
    # sketches whose hash densities differ 50x force several value-bounded phases
    rng = np.random.default_rng(5)
    s = 1500
    rows = []
    for i in range(48):
        scale = (1 << 60) // (1 + 49 * (i % 3 == 0))
        rows.append(np.unique(rng.integers(0, scale, size=s, dtype=np.uint64))[:s])
    h = np.zeros((48, s), dtype=np.uint64); sz = np.zeros(48, dtype=np.uint32)
    for i, r in enumerate(rows):
        h[i, :len(r)] = r; sz[i] = len(r)
    ln = np.full(48, 3_000_000, dtype=np.uint64)
    got, passed = dctx.dist_tile((h, sz, ln), (h[:20], sz[:20], ln[:20]), s, 21, 4.0 ** 21)
    _compare(got, passed, _oracle_matrix(oracle, (h, sz, ln), (h[:20], sz[:20], ln[:20]), s, 21, 4.0 ** 21))


def test_dist_literal_unsorted_fp_lists(ctx, oracle):
    # fp mode: unsorted, repeating 32-bit lists, k=1, kmerSpace=10
    rng = np.random.default_rng(6)
    n, s = 7, 2000
    h = rng.integers(0, 3000, size=(n, s)).astype(np.uint64)
    sz = np.full(n, s, dtype=np.uint32)
    ln = rng.integers(9000, 11000, size=n).astype(np.uint64)
    got, passed = ctx.dist_tile((h, sz, ln), (h, sz, ln), 1000, 1, 10.0, sorted_unique=False)
    _compare(got, passed, _oracle_matrix(oracle, (h, sz, ln), (h, sz, ln), 1000, 1, 10.0))
    # declared sorted but is not: the library must notice and still return the literal result
    got2, passed2 = ctx.dist_tile((h, sz, ln), (h, sz, ln), 1000, 1, 10.0, sorted_unique=True)
    assert np.array_equal(got2, got) and np.array_equal(passed2, passed)


def test_dist_empty_sketches(dctx, oracle):
    h = np.zeros((3, 10), dtype=np.uint64); h[1, :4] = [5, 9, 11, 40]; h[2, :2] = [9, 40]
    sz = np.array([0, 4, 2], dtype=np.uint32)
    ln = np.array([100, 200, 300], dtype=np.uint64)
    got, passed = dctx.dist_tile((h, sz, ln), (h, sz, ln), 10, 21, 4.0 ** 21)
    _compare(got, passed, _oracle_matrix(oracle, (h, sz, ln), (h, sz, ln), 10, 21, 4.0 ** 21))


def test_dist_large_sketches_many_phases(dctx, oracle):
    """s = 10000 (config 5's shape): every pair needs dozens of value-bounded phases."""
    rng = np.random.default_rng(10)
    s = 10000
    rh, rs = sorted_sketch_panel(rng, 20, s, n_clusters=3, shared=0.7)
    qh, qs = sorted_sketch_panel(rng, 9, s, n_clusters=3, shared=0.7)
    qh[0] = rh[0]; qs[0] = rs[0]
    rl = np.full(20, 5_000_000, dtype=np.uint64); ql = np.full(9, 4_000_000, dtype=np.uint64)
    got, passed = dctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 32, 4.0 ** 32)
    _compare(got, passed, _oracle_matrix(oracle, (rh, rs, rl), (qh, qs, ql), s, 32, 4.0 ** 32))


def test_dist_full_size_properties(ctx):
    """4000 x 4000 all-vs-all at s=1000: symmetry, identity diagonal, bounds -- properties that need no oracle."""
    rng = np.random.default_rng(11)
    n, s = 4000, 1000
    h, sz = sorted_sketch_panel(rng, n, s, n_clusters=20, shared=0.6, ragged=False)
    ln = np.full(n, 5_000_000, dtype=np.uint64)
    got, passed = ctx.dist_tile((h, sz, ln), (h, sz, ln), s, 21, 4.0 ** 21)
    assert np.array_equal(got["numer"], got["numer"].T) and np.array_equal(got["denom"], got["denom"].T)
    assert np.array_equal(got["distance"], got["distance"].T) and np.array_equal(got["pvalue"], got["pvalue"].T)
    assert (np.diag(got["numer"]) == s).all() and (np.diag(got["distance"]) == 0).all() and (got["denom"] == s).all()
    assert (got["numer"] <= got["denom"]).all() and passed.all()
    same = (np.arange(n)[:, None] % 20) == (np.arange(n)[None, :] % 20)
    assert got["numer"][same].min() > 100 and got["numer"][~same].max() < 10


def test_dist_rank32_equals_u64_kernel(ctx):
    """600 x 900 ragged panels incl. values near 2^64 and equal hashes across panels: both tile kernels byte-identical."""
    rng = np.random.default_rng(12)
    s = 1000
    rh, rs = sorted_sketch_panel(rng, 900, s, n_clusters=7, shared=0.7)
    qh, qs = sorted_sketch_panel(rng, 600, s, n_clusters=7, shared=0.7)
    qh[:40] = rh[100:140]; qs[:40] = rs[100:140]
    top = np.unique(rng.integers((1 << 64) - 5000, (1 << 64) - 2, size=700, dtype=np.uint64))   # high words all ones: the
    rh[5, :len(top)] = top; rs[5] = len(top)                                                     # u64 path must fall back
    rl = rng.integers(1000, 6_000_000, size=900).astype(np.uint64)
    ql = rng.integers(1000, 6_000_000, size=600).astype(np.uint64)
    got32, pass32 = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.3)
    ctx.set_dist_mode(force64=True)
    try:
        got64, pass64 = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.3)
    finally:
        ctx.set_dist_mode(force64=False)
    assert np.array_equal(got32, got64) and np.array_equal(pass32, pass64)


@pytest.mark.parametrize("n,ragged", [(64, False), (288, False), (300, True), (500, False), (1111, True)])
def test_dist_rank32_sizes_sweep(ctx, n, ragged):
    """All-vs-all at the panel sizes that exposed a miscompiled merge loop in the 32-bit kernel (hang / illegal
    address from 288 same-size sketches on): the 32-bit kernel must equal the 64-bit one bit for bit."""
    rng = np.random.default_rng(11)
    s = 1000
    h, sz = sorted_sketch_panel(rng, n, s, n_clusters=20, shared=0.6, ragged=ragged)
    ln = np.full(n, 5_000_000, dtype=np.uint64)
    got32, pass32 = ctx.dist_tile((h, sz, ln), (h, sz, ln), s, 21, 4.0 ** 21)
    ctx.set_dist_mode(force64=True)
    try:
        got64, pass64 = ctx.dist_tile((h, sz, ln), (h, sz, ln), s, 21, 4.0 ** 21)
    finally:
        ctx.set_dist_mode(force64=False)
    assert np.array_equal(got32, got64) and np.array_equal(pass32, pass64)
    full = sz == s
    assert (np.diag(got32["numer"])[full] == s).all()


def test_dist_structured_edge_cases(dctx, oracle):
    """Hand-built list shapes around the loop's exits (CommandDistance.cpp:376-400): sizes 0/1/s-1/s/s+1, identical,
    disjoint-interleaved, one list a prefix / suffix / subset of the other, all matches after the cut-off."""
    s = 64
    base = np.arange(1, 4 * s + 1, dtype=np.uint64) * np.uint64(1 << 40)
    lists = [
        base[:0], base[:1], base[:s - 1], base[:s], base[:s + 1],      # sizes around s
        base[::2][:s], base[1::2][:s],                                  # disjoint, perfectly interleaved
        base[s:2 * s], base[s // 2:s // 2 + s],                         # shifted windows (suffix/prefix overlaps)
        base[:2 * s:3], base[2 * s:3 * s],                              # sparse subset; entirely beyond the others
        np.concatenate([base[1:s:2], base[3 * s:3 * s + s // 2]]),      # shares only early elements
        np.concatenate([base[:3], base[3 * s + 5:3 * s + 5 + s - 3]]),  # three early matches, the rest far away
        base[s - 1:s], base[4 * s - 1:],                                # single elements at the cut-off / at the very end
    ]
    n = len(lists)
    width = max(len(x) for x in lists)
    h = np.zeros((n, max(width, 1)), dtype=np.uint64)
    sz = np.zeros(n, dtype=np.uint32)
    for i, x in enumerate(lists):
        h[i, :len(x)] = x
        sz[i] = len(x)
    ln = np.arange(1000, 1000 + n, dtype=np.uint64) * np.uint64(977)
    for s_cmp in (s, s - 1, 1, 3 * s):
        got, passed = dctx.dist_tile((h, sz, ln), (h, sz, ln), s_cmp, 21, 4.0 ** 21)
        _compare(got, passed, _oracle_matrix(oracle, (h, sz, ln), (h, sz, ln), s_cmp, 21, 4.0 ** 21))


def test_dist_small_universe_many_ties(ctx, oracle):
    """Sketches drawn from a universe of only 4000 hash values: most steps of most merges are ties (both lists advance),
    columns run out at very different rows, many pairs end by exhaustion instead of reaching s."""
    rng = np.random.default_rng(13)
    universe = np.sort(rng.choice(1 << 62, size=4000, replace=False).astype(np.uint64))
    n, width = 230, 1800
    h = np.zeros((n, width), dtype=np.uint64)
    sz = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        m = int(rng.integers(0, width + 1)) if i % 5 else width
        pick = np.sort(rng.choice(4000, size=m, replace=False))
        h[i, :m] = universe[pick]
        sz[i] = m
    ln = rng.integers(10_000, 9_000_000, size=n).astype(np.uint64)
    for s_cmp in (1000, 1800, 3000):
        got32, pass32 = ctx.dist_tile((h, sz, ln), (h[:70], sz[:70], ln[:70]), s_cmp, 21, 4.0 ** 21)
        ctx.set_dist_mode(force64=True)
        try:
            got64, pass64 = ctx.dist_tile((h, sz, ln), (h[:70], sz[:70], ln[:70]), s_cmp, 21, 4.0 ** 21)
        finally:
            ctx.set_dist_mode(force64=False)
        assert np.array_equal(got32, got64) and np.array_equal(pass32, pass64)
        # a sample of pairs against the oracle's literal loop
        for q, r in [(0, 0), (1, 5), (7, 229), (33, 100), (69, 64), (12, 13), (5, 1)]:
            w = oracle.compare(h[r, :sz[r]], h[q, :sz[q]], int(ln[r]), int(ln[q]), s_cmp, 21, 4.0 ** 21, 1.0, 1.0)
            g = got32[q, r]
            assert int(g["numer"]) == w["numer"] and int(g["denom"]) == w["denom"], (s_cmp, q, r, g, w)
            assert g["distance"] == pytest.approx(w["distance"], rel=RTOL, abs=0)


@pytest.mark.parametrize("n_r,n_q", [(700, 333), (300, 261), (1025, 513)])
def test_dist_device_path_grouped_equals_ungrouped(ctx, n_r, n_q):
    """The device entry point (results stay in HBM) reorders both panels so that related sketches share tiles; the output
    must not depend on it: grouped+pruned = pruned = every pair merged = host path, byte for byte."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(14)
    s = 500
    rh, rs = sorted_sketch_panel(rng, n_r, s, n_clusters=9, shared=0.6)
    qh, qs = sorted_sketch_panel(rng, n_q, s, n_clusters=9, shared=0.6)
    qh[:50] = rh[200:250]; qs[:50] = rs[200:250]
    qh[60:80] = 0; qs[60:80] = 0                                   # empty query sketches in the middle
    rl = rng.integers(1000, 6_000_000, size=n_r).astype(np.uint64)
    ql = rng.integers(1000, 6_000_000, size=n_q).astype(np.uint64)
    want, _ = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, raw=True)
    dev = torch.device("cuda", 0)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt)).to(dev)
    d_rh, d_rs, d_rl = t(rh, np.int64), t(rs.astype(np.uint32), np.int32), t(rl, np.int64)
    d_qh, d_qs, d_ql = t(qh, np.int64), t(qs.astype(np.uint32), np.int32), t(ql, np.int64)
    outs = []
    # default = regroup only when the marked pairs are scattered; group=True always; no_group: neither regrouping nor the tile list
    for kw in (dict(), dict(group=True), dict(no_group=True), dict(no_prune=True)):
        ctx.set_dist_mode(**kw)
        try:
            out = torch.zeros(n_q * n_r * 24, dtype=torch.uint8, device=dev)
            ctx.dist_tile_dev((d_rh.data_ptr(), d_rs.data_ptr(), d_rl.data_ptr(), n_r, s), (d_qh.data_ptr(), d_qs.data_ptr(), d_ql.data_ptr(), n_q, s),
                              s, 21, 4.0 ** 21, out.data_ptr())
            torch.cuda.synchronize()
            outs.append(out.cpu().numpy().tobytes())
        finally:
            ctx.set_dist_mode()
    assert outs[0] == outs[1] == outs[2] == outs[3]
    assert outs[0] == np.ascontiguousarray(want).tobytes()


# ---- fpm_dist_hits: only the pairs `mash dist -d D -v P` prints, in its output order ---------------------------------

def _hits_from_matrix(mat_raw):
    """What fpm_dist_hits must return, derived from the raw fpm_dist_tile matrix of the same call."""
    q, r = np.nonzero((mat_raw["denom"] & 0x80000000) != 0)        # row-major: sorted by (query, ref)
    sel = mat_raw[q, r]
    return q, r, sel


def _clustered_case(seed, n_r, n_q, s):
    rng = np.random.default_rng(seed)
    h, sz = sorted_sketch_panel(rng, n_r + n_q, s, n_clusters=9, shared=0.6)     # one family tree, split into the two panels
    rh, rs, qh, qs = h[:n_r].copy(), sz[:n_r].copy(), h[n_r:].copy(), sz[n_r:].copy()
    c = min(n_q, n_r) // 4
    qh[:c] = rh[c:2 * c]; qs[:c] = rs[c:2 * c]                    # identical pairs
    qh[c + 2:c + 5] = 0; qs[c + 2:c + 5] = 0                        # empty query sketches
    rh[1] = 0; rs[1] = 0; rh[n_r - 1] = 0; rs[n_r - 1] = 0          # and empty references: empty x empty has distance 0 without sharing a hash
    rl = rng.integers(1000, 6_000_000, size=n_r).astype(np.uint64)
    ql = rng.integers(1000, 6_000_000, size=n_q).astype(np.uint64)
    return (rh, rs, rl), (qh, qs, ql)


@pytest.mark.parametrize("max_d,max_p", [(0.1, 1.0), (1.0, 1e-10), (0.05, 1e-3), (1.0, 1.0), (0.0, 1.0)])
def test_dist_hits_equal_the_passing_pairs_of_the_matrix(dctx, max_d, max_p):
    s = 400
    ref, qry = _clustered_case(21, 300, 131, s)
    mat, _ = dctx.dist_tile(ref, qry, s, 21, 4.0 ** 21, max_distance=max_d, max_pvalue=max_p, raw=True)
    q, r, sel = _hits_from_matrix(mat)
    hits = dctx.dist_hits(ref, qry, s, 21, 4.0 ** 21, max_distance=max_d, max_pvalue=max_p, capacity=7)   # forces the retry
    assert len(hits) == len(q)
    if max_d == 1.0 and max_p == 1.0:
        assert len(hits) == 300 * 131                                 # no filter: every pair, unmarked ones included
    assert np.array_equal(hits["query"], q) and np.array_equal(hits["ref"], r)
    assert np.array_equal(hits["numer"], sel["numer"]) and np.array_equal(hits["denom"], sel["denom"] & 0x7fffffff)
    assert hits["distance"].tobytes() == sel["distance"].tobytes() and hits["pvalue"].tobytes() == sel["pvalue"].tobytes()


def test_dist_hits_against_the_oracle(ctx, oracle):
    s = 200
    ref, qry = _clustered_case(5, 40, 23, s)
    hits = ctx.dist_hits(ref, qry, s, 21, 4.0 ** 21, max_distance=0.06, max_pvalue=1e-3)
    want = _oracle_matrix(oracle, ref, qry, s, 21, 4.0 ** 21, 0.06, 1e-3)
    exp = [(q, r, w) for q, row in enumerate(want) for r, w in enumerate(row) if w["passed"]]
    assert 20 < len(exp) < 23 * 40 and len(hits) == len(exp)
    for h, (q, r, w) in zip(hits, exp):
        assert (int(h["query"]), int(h["ref"]), int(h["numer"]), int(h["denom"])) == (q, r, w["numer"], w["denom"])
        assert h["distance"] == pytest.approx(w["distance"], rel=RTOL, abs=0)
        assert h["pvalue"] == pytest.approx(w["pvalue"], rel=RTOL, abs=1e-300)


def test_dist_hits_literal_path_and_capacity_error(ctx):
    import fpmash_b200 as fpm
    rng = np.random.default_rng(8)
    n_r, n_q, s = 50, 31, 64
    rh = rng.integers(0, 300, size=(n_r, s)).astype(np.uint64)        # unsorted, repeating: fp-mode lists
    qh = rng.integers(0, 300, size=(n_q, s)).astype(np.uint64)
    rs = np.full(n_r, s, dtype=np.uint32); qs = np.full(n_q, s, dtype=np.uint32)
    rl = np.full(n_r, 5000, dtype=np.uint64); ql = np.full(n_q, 5000, dtype=np.uint64)
    mat, _ = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.3, sorted_unique=False, raw=True)
    q, r, sel = _hits_from_matrix(mat)
    hits = ctx.dist_hits((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.3, sorted_unique=False)
    assert len(q) > 0 and np.array_equal(hits["query"], q) and np.array_equal(hits["ref"], r)
    assert hits["distance"].tobytes() == sel["distance"].tobytes()
    small = np.zeros(max(1, len(q) // 2), dtype=fpm.HIT_DTYPE)
    with pytest.raises(fpm.FpmError) as e:
        ctx.dist_hits((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.3, sorted_unique=False, out=small)
    assert e.value.code == fpm.FPM_ERR_CAPACITY and str(len(q)) in str(e.value)
    # nothing passes / empty panels
    mat0, _ = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.0, max_pvalue=0.0, sorted_unique=False, raw=True)
    none = ctx.dist_hits((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.0, max_pvalue=0.0, sorted_unique=False)
    assert len(none) == len(_hits_from_matrix(mat0)[0]) == 0
    empty = ctx.dist_hits((rh[:0], rs[:0], rl[:0]), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.3)
    assert len(empty) == 0


def test_dist_hits_device_entry_point(ctx):
    torch = pytest.importorskip("torch")
    import fpmash_b200 as fpm
    s = 500
    ref, qry = _clustered_case(33, 700, 333, s)
    (rh, rs, rl), (qh, qs, ql) = ref, qry
    mat, _ = ctx.dist_tile(ref, qry, s, 21, 4.0 ** 21, max_distance=0.15, raw=True)
    q, r, sel = _hits_from_matrix(mat)
    dev = torch.device("cuda", 0)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a).view(dt)).to(dev)
    d_rh, d_rs, d_rl = t(rh, np.int64), t(rs.astype(np.uint32), np.int32), t(rl, np.int64)
    d_qh, d_qs, d_ql = t(qh, np.int64), t(qs.astype(np.uint32), np.int32), t(ql, np.int64)
    cap = len(q) + 5
    steps = torch.zeros(1, dtype=torch.int64, device=dev)
    for kw in (dict(), dict(group=True), dict(no_group=True), dict(no_prune=True)):
        ctx.set_dist_mode(**kw)
        try:
            out = torch.zeros(cap * 32, dtype=torch.uint8, device=dev)
            n = ctx.dist_hits_dev((d_rh.data_ptr(), d_rs.data_ptr(), d_rl.data_ptr(), 700, s), (d_qh.data_ptr(), d_qs.data_ptr(), d_ql.data_ptr(), 333, s),
                                  s, 21, 4.0 ** 21, out.data_ptr(), cap, d_steps_ptr=steps.data_ptr(), max_distance=0.15)
        finally:
            ctx.set_dist_mode()
        hits = out.cpu().numpy().view(fpm.HIT_DTYPE)[:n]
        assert n == len(q) and np.array_equal(hits["query"], q) and np.array_equal(hits["ref"], r)
        assert np.array_equal(hits["numer"], sel["numer"]) and hits["distance"].tobytes() == sel["distance"].tobytes()
        assert hits["pvalue"].tobytes() == sel["pvalue"].tobytes()
    assert int(steps.item()) > 0


def test_dist_pruning_chained_components(ctx, oracle):
    """The marking pass stops a query once it has marked as many references as its connected component holds.  Chains
    (q shares with r0 only, r0 with r1, r1 with r2 ...) put many references in the component that the query never
    reaches: the walk must not stop early and must not mark them; cliques are the case where it does stop."""
    rng = np.random.default_rng(99)
    s, n = 64, 96
    base = np.sort(rng.choice(1 << 40, size=(n + 1) * s, replace=False).astype(np.uint64)).reshape(n + 1, s)
    rh = np.zeros((n, s), dtype=np.uint64)
    for i in range(n):                                   # reference i = half of block i + half of block i+1: a chain
        rh[i] = np.sort(np.concatenate([base[i, ::2], base[i + 1, 1::2]]))
    rh[40:60] = np.sort(np.concatenate([np.tile(base[40, :8], (20, 1)), rng.integers(1 << 41, 1 << 42, size=(20, s - 8)).astype(np.uint64)], axis=1), axis=1)  # a clique
    qh = np.zeros((n, s), dtype=np.uint64)
    qh[:] = base[:n]                                     # query i = block i: shares with references i-1 and i only
    qh[45] = rh[45]
    rs = np.full(n, s, dtype=np.uint32); qs = np.full(n, s, dtype=np.uint32)
    rl = np.full(n, 100_000, dtype=np.uint64); ql = np.full(n, 100_000, dtype=np.uint64)
    ctx.set_dist_mode(saturate=True)                     # (small panels would otherwise skip the component bound)
    try:
        got, passed = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21)
        _compare(got, passed, _oracle_matrix(oracle, (rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21))
        assert int((got["numer"] > 0).sum()) > 2 * n - 40
        hits = ctx.dist_hits((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, max_distance=0.9)
    finally:
        ctx.set_dist_mode()
    assert len(hits) == int((got["distance"] <= 0.9).sum())


def test_resident_reference_panel_equals_one_call(dctx, fpm):
    """fpm_dist_set_reference: the reference panel is uploaded and indexed once; query chunks of any size, in any order, give
    the rows of the single full call (and an explicit panel afterwards replaces the resident one)."""
    rng = np.random.default_rng(321)
    s, n_ref, n_qry = 128, 300, 517
    rh, rs = sorted_sketch_panel(rng, n_ref, s, n_clusters=5)
    qh, qs = sorted_sketch_panel(rng, n_qry, s, n_clusters=5)
    qh[:4] = rh[:4]; qs[:4] = rs[:4]
    rl = rng.integers(1000, 6_000_000, size=n_ref).astype(np.uint64)
    ql = rng.integers(1000, 6_000_000, size=n_qry).astype(np.uint64)
    want, _ = dctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, raw=True)
    dctx.dist_set_reference((rh, rs, rl))
    try:
        for lo, hi in ((400, 517), (0, 1), (1, 257), (257, 400), (0, 517)):
            got, _ = dctx.dist_tile(None, (qh[lo:hi], qs[lo:hi], ql[lo:hi]), s, 21, 4.0 ** 21, raw=True)
            assert got.tobytes() == want[lo:hi].tobytes(), (lo, hi)
        # another reference panel through the ordinary call, then the resident one is gone
        other, _ = dctx.dist_tile((qh[:50], qs[:50], ql[:50]), (qh[:20], qs[:20], ql[:20]), s, 21, 4.0 ** 21, raw=True)
        assert other.shape == (20, 50)
        with pytest.raises(fpm.FpmError):
            dctx.dist_tile(None, (qh[:5], qs[:5], ql[:5]), s, 21, 4.0 ** 21)
    finally:
        dctx.dist_set_reference(None)


def test_dist_long_reference_chains_same_in_every_mode(ctx, fpm):
    """the component pass must never make a query stop marking early.  Long chains of references (reference i shares
    a few hashes with i+1 only: trees 3000 deep while the union-find forms) plus queries hanging off them at random places;
    the result with the component bound forced on must equal the result without it, without grouping and without pruning."""
    rng = np.random.default_rng(4242)
    s, n_r, n_q = 48, 3000, 700
    link = np.sort(rng.choice(1 << 44, size=(n_r + 1) * 6, replace=False).astype(np.uint64)).reshape(n_r + 1, 6)
    priv = rng.integers(1 << 45, 1 << 46, size=(n_r, s - 12)).astype(np.uint64)
    rh = np.sort(np.concatenate([link[:-1], link[1:], priv], axis=1), axis=1)           # reference i: links i and i+1
    qh = np.zeros((n_q, s), dtype=np.uint64)
    at = rng.integers(0, n_r, size=n_q)
    for q in range(n_q):
        own = rng.integers(1 << 47, 1 << 48, size=s - 6).astype(np.uint64)
        qh[q] = np.sort(np.concatenate([link[at[q]], own]))                             # touches references at[q]-1 and at[q]
    for a in (rh, qh):
        assert (np.diff(a.astype(np.float64), axis=1) > 0).all()
    rs = np.full(n_r, s, dtype=np.uint32); qs = np.full(n_q, s, dtype=np.uint32)
    rl = np.full(n_r, 1_000_000, dtype=np.uint64); ql = np.full(n_q, 1_000_000, dtype=np.uint64)
    outs = []
    for mode in ({"saturate": True}, {}, {"group": True}, {"no_group": True}, {"no_prune": True}):
        ctx.set_dist_mode(**mode)
        try:
            for _ in range(3 if mode.get("saturate") else 1):                            # the race was rare: repeat
                got, _ = ctx.dist_tile((rh, rs, rl), (qh, qs, ql), s, 21, 4.0 ** 21, raw=True)
                outs.append(got.tobytes())
        finally:
            ctx.set_dist_mode()
    assert all(o == outs[0] for o in outs)
    got = np.frombuffer(outs[0], dtype=fpm.PAIR_DTYPE).reshape(n_q, n_r)
    assert int((got["numer"] == 6).sum()) >= n_q                                          # every query found its neighbours
