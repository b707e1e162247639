"""Parameters of the deterministic synthetic inputs: vocabulary and Zipf CDF tables
consumed by the device generator (csrc/ops.cu: dampr_synth_text). Bench / test tooling — not on any
product path. oracle/gen.py holds the numpy restatement of the generator and uses the same tables."""
import numpy as np


def make_vocab(V, seed=1234):
    """(vocab_bytes uint8[], vocab_off uint32[V+1]) — V lowercase ASCII words, lengths U[2,11]."""
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
