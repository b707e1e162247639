"""Host-side decoding of the 64-bit token key codes produced by text_count_kernel (csrc/text.cu).

Short tokens carry an exact, injective code (no hashing, so no collision can merge two words):
  * TOK_WS: up to 9 bytes, 7 bits each, first byte in the low bits;
  * TOK_NONWORD_LOWER*: up to 12 symbols base 38 ('0'-'9' -> 1..10, '_' -> 11, 'a'-'z' -> 12..37),
    first symbol least significant.
Longer tokens carry (1 << 63) | hash and a representative (offset << 20 | length) from which the
bytes are read back; dampr_text_verify byte-compares every such token against its representative.
"""
import numpy as np

from . import device as dev

_SYM = np.frombuffer(b"\x000123456789_abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)
HASHED_BIT = np.uint64(1 << 63)


def decode_exact(codes, mode):
    """numpy uint64 codes (bit 63 clear) -> list of str."""
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    n = len(codes)
    if n == 0:
        return []
    if mode == dev.TOK_WS:
        width = 9
        out = np.zeros((n, width), dtype=np.uint8)
        c = codes.copy()
        for i in range(width):
            out[:, i] = (c & np.uint64(127)).astype(np.uint8)
            c >>= np.uint64(7)
    else:
        width = 12
        out = np.zeros((n, width), dtype=np.uint8)
        c = codes.copy()
        for i in range(width):
            out[:, i] = _SYM[(c % np.uint64(38)).astype(np.int64)]
            c //= np.uint64(38)
    raw = out.view("S%d" % width).ravel()
    return [b.decode("ascii") for b in raw.tolist()]


def decode_table(codes, reps, mode, read_bytes):
    """codes/reps from Table.fetch -> list of str. read_bytes(offset, length) -> bytes supplies the
    text of hashed (long) tokens; they are lower-cased in the NONWORD modes like the device does."""
    codes = np.ascontiguousarray(codes, dtype=np.uint64)
    hashed = (codes & HASHED_BIT) != 0
    words = [None] * len(codes)
    ex_idx = np.flatnonzero(~hashed)
    for i, w in zip(ex_idx.tolist(), decode_exact(codes[ex_idx], mode)):
        words[i] = w
    for i in np.flatnonzero(hashed).tolist():
        rep = int(reps[i])
        b = bytes(read_bytes(rep >> 20, rep & 0xFFFFF))
        w = b.decode("ascii")
        words[i] = w if mode == dev.TOK_WS else w.lower()
    return words
