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


from dampr_b200.synth import make_vocab, make_cdf  # noqa: E402,F401  (same tables as the device generator)


def pad_tail(total):
    """bytes appended in place of the final newline so that the size is a multiple of 64."""
    pad = (64 - (total % 64)) % 64
    if pad == 1:
        pad = 65
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


def text(seed, n_lines, V=50000, s=1.1, vocab=None, cdf=None, pad=True):
    """Synthetic corpus as bytes (numpy implementation; fine up to a few hundred MB)."""
    if vocab is None:
        vocab = make_vocab(V)
    if cdf is None:
        cdf = make_cdf(V, s)
    vbytes, voff = vocab
    V = len(voff) - 1
    if n_lines == 0:
        return b""
    out_parts = []
    total = 0
    BLOCK = 1 << 18
    for lo in range(0, n_lines, BLOCK):
        hi = min(n_lines, lo + BLOCK)
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
        out_parts.append(out.tobytes())
        total += nbytes
    data = b"".join(out_parts)
    if pad:
        tail, _ = pad_tail(len(data))
        if tail:
            data = data[:-1] + tail
    return data


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
