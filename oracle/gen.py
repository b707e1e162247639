"""Deterministic synthetic inputs — TEST INFRASTRUCTURE, not product code.

Integer-only algorithms so that the numpy implementation here and the CUDA generator
(dampr_b200/csrc/ops.cu: synth_len_kernel / synth_write_kernel / synth_kv_kernel) produce
byte-identical data:

text(seed, n_lines, V, s): vocabulary of V random lowercase words of length U[2,11]
  (numpy default_rng(seed)), Zipf(s) ranks through a 64-bit integer CDF, line i has
  5 + splitmix64(seed ^ i*GOLD) % 15 tokens, token j = searchsorted(cdf, splitmix64(r0 + (j+1)*C2)),
  tokens joined by one space, '\n' terminated, ASCII only, total padded to a multiple of 64 bytes by
  extending the last line with the words "a"/"aa".
kv(seed, n, n_keys): key = (splitmix64(seed+i) % n_keys) * GOLD mod 2^64, value in [-1000, 1000).
"""
import numpy as np

GOLD = np.uint64(0x9E3779B97F4A7C15)
C2 = np.uint64(0xD1B54A32D192ED03)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
KVX = np.uint64(0x5851F42D4C957F2D)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x + GOLD
        x = (x ^ (x >> np.uint64(30))) * M1
        x = (x ^ (x >> np.uint64(27))) * M2
        return x ^ (x >> np.uint64(31))


def make_vocab(V, seed=1234):
    """(vocab_bytes uint8[], vocab_off uint32[V+1]) — V lowercase ASCII words, lengths U[2,11].
    The product's bench tooling (dampr_b200/synth.py) builds the same tables for the device generator;
    tests/test_oracle_golden.py pins the two against each other. The oracle imports nothing of the product."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(2, 12, size=V).astype(np.uint32)
    off = np.zeros(V + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    letters = rng.integers(0, 26, size=int(off[-1])).astype(np.uint8) + np.uint8(ord("a"))
    return letters, off


def make_cdf(V, s=1.1):
    """uint64[V]: cdf[i] = floor(2^64 * P(rank <= i+1)) for Zipf(s), last entry 2^64-1."""
    p = np.arange(1, V + 1, dtype=np.float64) ** (-float(s))
    c = np.cumsum(p)
    c /= c[-1]
    scaled = np.minimum(np.floor(c * 18446744073709551616.0), 18446744073709549568.0)
    cdf = scaled.astype(np.uint64)
    cdf[-1] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return cdf


def pad_tail(total, mult=64):
    """bytes appended in place of the final newline so that the size is a multiple of `mult` (64 in the
    generator; the bench's CPU arms use lcm(64, cpu count) so that the reference's st_size / cpu_count
    chunk size is integral)."""
    pad = (mult - (total % mult)) % mult
    if pad == 1:
        pad = mult + 1
    if pad == 0 or total == 0:
        return b"", 0
    tail = b""
    left = pad
    while left > 0:
        if left == 3:
            tail += b" aa"
            left -= 3
        else:
            tail += b" a"
            left -= 2
    return tail + b"\n", pad


BLOCK = 1 << 18
_POOL_STATE = None


def _block(seed, lo, hi, vbytes, voff, cdf):
    """bytes of lines [lo, hi) (every line '\n' terminated)"""
    V = len(voff) - 1
    if True:
        i = np.arange(lo, hi, dtype=np.uint64)
        with np.errstate(over="ignore"):
            r0 = splitmix64(np.uint64(seed) ^ (i * GOLD))
        ntok = (5 + (r0 % np.uint64(15))).astype(np.int64)
        T = int(ntok.sum())
        line_of = np.repeat(np.arange(hi - lo), ntok)
        first = np.zeros(hi - lo, dtype=np.int64)
        np.cumsum(ntok[:-1], out=first[1:])
        j = np.arange(T, dtype=np.int64) - first[line_of]
        with np.errstate(over="ignore"):
            r = splitmix64(r0[line_of] + (j.astype(np.uint64) + np.uint64(1)) * C2)
        w = np.searchsorted(cdf, r, side="left")
        w = np.minimum(w, V - 1)
        wl = (voff[w + 1] - voff[w]).astype(np.int64)
        tok_bytes = wl + 1
        tok_out = np.zeros(T, dtype=np.int64)
        np.cumsum(tok_bytes[:-1], out=tok_out[1:])
        nbytes = int(tok_bytes.sum())
        out = np.empty(nbytes, dtype=np.uint8)
        # separators
        sep = np.full(T, ord(" "), dtype=np.uint8)
        last = first + ntok - 1
        sep[last] = ord("\n")
        out[tok_out + wl] = sep
        # word bytes: ragged gather
        W = int(wl.sum())
        tok_of = np.repeat(np.arange(T), wl)
        wfirst = np.zeros(T, dtype=np.int64)
        np.cumsum(wl[:-1], out=wfirst[1:])
        k = np.arange(W, dtype=np.int64) - wfirst[tok_of]
        out[tok_out[tok_of] + k] = vbytes[voff[w][tok_of].astype(np.int64) + k]
        return out.tobytes()


def text(seed, n_lines, V=50000, s=1.1, vocab=None, cdf=None, pad=True):
    """Synthetic corpus as bytes (numpy implementation; fine up to a few hundred MB)."""
    if vocab is None:
        vocab = make_vocab(V)
    if cdf is None:
        cdf = make_cdf(V, s)
    vbytes, voff = vocab
    if n_lines == 0:
        return b""
    data = b"".join(_block(seed, lo, min(n_lines, lo + BLOCK), vbytes, voff, cdf) for lo in range(0, n_lines, BLOCK))
    if pad:
        tail, _ = pad_tail(len(data))
        if tail:
            data = data[:-1] + tail
    return data


def _pool_block(rng):
    seed, vbytes, voff, cdf = _POOL_STATE
    return _block(seed, rng[0], rng[1], vbytes, voff, cdf)


def text_to_file(path, seed, n_lines, V=50000, s=1.1, procs=None, pad=True):
    """The same corpus written to `path`, blocks of lines generated by forked workers (the bench's CPU arms
    need ~1 GB samples on the GPU box's many host cores). Returns the byte count."""
    import multiprocessing as mp
    global _POOL_STATE
    vocab, cdf = make_vocab(V), make_cdf(V, s)
    _POOL_STATE = (seed, vocab[0], vocab[1], cdf)
    ranges = [(lo, min(n_lines, lo + BLOCK)) for lo in range(0, n_lines, BLOCK)]
    procs = max(1, min(procs or (mp.cpu_count() or 1), len(ranges)))
    total = 0
    with open(path, "wb") as f:
        if procs == 1:
            it = map(_pool_block, ranges)
            pool = None
        else:
            pool = mp.get_context("fork").Pool(procs)
            it = pool.imap(_pool_block, ranges)
        try:
            for blk in it:
                f.write(blk)
                total += len(blk)
        finally:
            if pool is not None:
                pool.close()
                pool.join()
        if pad and total:
            tail, _ = pad_tail(total)
            if tail:
                f.seek(total - 1)
                f.write(tail)
                total += len(tail) - 1
    _POOL_STATE = None
    return total


def kv(seed, n, n_keys):
    """(keys uint64[n], vals int64[n])."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        r = splitmix64(np.uint64(seed) + i)
        keys = (r % np.uint64(n_keys)) * GOLD
        r2 = splitmix64(r ^ KVX)
    vals = (r2 % np.uint64(2000)).astype(np.int64) - 1000
    return keys, vals


DIRTY_LINES = [
    b"", b" ", b"...", b"Hello, World!", b"  leading and trailing  ", b"it's a dog-eat-dog world_2",
    b"MiXeD CaSe mixed case MIXED", b"tab\tseparated\x0bvertical\x0cformfeed", b"under_score __init__ _x_",
    b"digits 123 4567 89a a89", b"a", b"a a a a a a", b"end.", b".start", b"x" * 13, b"y" * 12,
    b"abcdefghijklm abcdefghijklm ABCDEFGHIJKLM", b"\x1c\x1d\x1e\x1f control separators", b"~!@#$%^&*()",
    b"word " * 40, b"supercalifragilisticexpialidocious", b"aaaaaaaaaa bbbbbbbbb cccccccccc",
]


def dirty_text(seed=7, n_lines=5000, long_line_bytes=3000):
    """Correctness corpus: mixed case, punctuation runs, empty lines, long tokens, a long line."""
    rng = np.random.default_rng(seed)
    lines = []
    for _ in range(n_lines):
        k = int(rng.integers(0, len(DIRTY_LINES)))
        reps = int(rng.integers(1, 3))
        lines.append(b" ".join([DIRTY_LINES[k]] * reps) if DIRTY_LINES[k] else b"")
    lines.insert(n_lines // 2, (b"long line token " * (long_line_bytes // 16))[:long_line_bytes])
    return b"\n".join(lines) + b"\n"
