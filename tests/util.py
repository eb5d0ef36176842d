"""Seeded synthetic inputs shared by the tests."""
import numpy as np


def random_dna(rng, n, gc=0.5):
    p = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.choice(4, size=n, p=p)].tobytes()


def dirty_dna(rng, n, n_rate=0.01, lower_rate=0.2, iupac_rate=0.002):
    """DNA with N runs, lower case and IUPAC codes (window-validity edge cases)."""
    s = bytearray(random_dna(rng, n))
    for i in np.nonzero(rng.random(n) < lower_rate)[0]:
        s[i] = s[i] | 0x20
    for i in np.nonzero(rng.random(n) < n_rate)[0]:
        run = int(rng.integers(1, 6))
        s[i:i + run] = b"N" * len(s[i:i + run])
    for i in np.nonzero(rng.random(n) < iupac_rate)[0]:
        s[i] = rng.choice(list(b"RYKMSWBDHVnryu-*"))
    return bytes(s)


def mutate(rng, seq, rate):
    s = np.frombuffer(seq, dtype=np.uint8).copy()
    idx = np.nonzero(rng.random(len(s)) < rate)[0]
    s[idx] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=len(idx))]
    return s.tobytes()


def sorted_sketch_panel(rng, n, s, n_clusters=4, shared=0.5, bits=64, ragged=True):
    """n sorted duplicate-free sketches; members of a cluster share ~`shared` of their hashes."""
    hi = (1 << bits) - 1 if bits < 64 else (1 << 64) - 1
    scale = hi // 64
    cores = [np.unique(rng.integers(0, scale, size=2 * s, dtype=np.uint64)) for _ in range(n_clusters)]
    hashes = np.zeros((n, s), dtype=np.uint64)
    sizes = np.zeros(n, dtype=np.uint32)
    for i in range(n):
        core = cores[i % n_clusters]
        take = core[rng.random(len(core)) < shared]
        priv = rng.integers(0, scale, size=s, dtype=np.uint64)
        u = np.unique(np.concatenate([take, priv]))
        m = s if not ragged or rng.random() < 0.6 else int(rng.integers(0, s + 1))
        u = u[:m]
        hashes[i, :len(u)] = u
        sizes[i] = len(u)
    return hashes, sizes
