"""numpy mirror of the integer kinds of the device generator (csrc/gen.cu): element j of a
column depends on (seed, j) only.  Used to build host-resident inputs (the e2e path starts
from host buffers, like the reference after fread, npj.cpp:1036-1039) and to cross-check the
device generator.  Semantics follow generate_data_for_join (cpra2.cpp:1578-1696): distinct
non-zero keys, foreign keys = every key once then uniform picks, shuffled, payload = key*factor."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)
INNER_FACTOR = 0x6587F97D   # odd payload factors (the survey recovered factors of this shape)
OUTER_FACTOR = 0xDF56B8FB


def mix32(x):
    x = np.asarray(x, dtype=np.uint64) & M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & M32
    x ^= x >> np.uint64(16)
    return x


def _mix32_scalar(x):
    return int(mix32(np.array([x], np.uint64))[0])


def key_of_rank(rank, seed):
    salt = _mix32_scalar((seed * 2 + 1) & 0xFFFFFFFF)
    k = mix32((np.asarray(rank, np.uint64) + np.uint64(1)) ^ np.uint64(salt))
    repl = _mix32_scalar(salt) or 1
    return np.where(k == 0, np.uint64(repl), k).astype(np.uint32)


def permute_index(j, total, seed):
    bits = 1
    while (1 << bits) < total:
        bits += 1
    mask = np.uint64((1 << bits) - 1)
    sh = np.uint64((bits + 1) // 2)
    cs = [_mix32_scalar(seed ^ c) for c in (0x11111111, 0x22222222, 0x33333333, 0x44444444, 0x55555555, 0x66666666)]
    a = [np.uint64((cs[i] << 1) | 1) for i in (0, 2, 4)]
    c = [np.uint64(cs[i]) for i in (1, 3, 5)]

    def step(x):
        with np.errstate(over="ignore"):
            for r in range(3):
                x = (x * a[r] + c[r]) & mask
                x ^= x >> sh
        return x
    x = step(np.asarray(j, dtype=np.uint64).copy())
    bad = x >= np.uint64(total)
    while bad.any():
        x[bad] = step(x[bad])
        bad = x >= np.uint64(total)
    return x


def _chunks(n, step=1 << 24):
    for b in range(0, n, step):
        yield b, min(n, b + step)


def generate(kind, tuples, domain, seed, order_seed, payload_factor, first=0, total=None, out=None):
    """kind 0: unique keys (a seeded permutation of ranks [0,total)); kind 1: foreign keys into
    `domain` build keys.  Returns (keys, vals) uint32; `out` may supply the two arrays (e.g.
    pinned memory)."""
    total = tuples if total is None else total
    keys, vals = out if out is not None else (np.empty(tuples, np.uint32), np.empty(tuples, np.uint32))
    for b, e in _chunks(tuples):
        j = np.arange(first + b, first + e, dtype=np.uint64)
        t = permute_index(j, total, order_seed)
        if kind == 0:
            rank = t
        elif kind == 1:
            h = mix32((t & M32) ^ mix32(((t >> np.uint64(32)) + np.uint64(order_seed)) & M32))
            pick = (h * np.uint64(domain)) >> np.uint64(32)
            rank = np.where(t < np.uint64(domain), t, pick)
        else:
            raise ValueError("the skewed kind exists on the device only (double-precision inversion)")
        k = key_of_rank(rank, seed)
        keys[b:e] = k
        with np.errstate(over="ignore"):
            vals[b:e] = (k.astype(np.uint64) * np.uint64(payload_factor) & M32).astype(np.uint32)
    return keys, vals


def workload(name, scale=1.0, seed=42):
    """Sizes of BASELINE.json's configs: name -> (|R|, kind_R, |S|, kind_S)."""
    cfg = {
        "npj_cfg1": (1 << 24, 1 << 28, 1),     # 16M unique x 256M foreign keys
        "phj_cfg2": (1 << 27, 1 << 27, 0),     # 128M x 128M, both permutations of one key set
        "small_cfg3": (1 << 16, 1 << 30, 1),   # 64K x 1B foreign keys
    }[name]
    nr, ns, ks = cfg
    return max(1, int(nr * scale)), max(1, int(ns * scale)), ks
