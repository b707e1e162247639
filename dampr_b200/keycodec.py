"""Python keys -> 64-bit device sort codes (host side of the shuffle for non-lowered stages).

The device sorts/groups 16-byte (code, row id) records; this module decides the code:

  INT    all keys are ints in [-2^63, 2^63): code = key + 2^63. Exact, order preserving.
  FLOAT  floats (or floats mixed with ints |k| <= 2^53): IEEE bits made monotone. Exact, ordered.
  STR    str keys: first 8 UTF-8 bytes, big endian. Ordered; NOT exact (keys sharing an 8-byte
         prefix share a code) — the host refines runs of equal codes by comparing the real keys.
  BYTES  same for bytes.
  HASH   anything else (None, tuples, mixed types): blake2b-64 of a canonical form. Not ordered,
         not exact; runs of equal codes are split by real-key equality.
  TUPLE  (sort_by only) tuples whose first elements are all INT/FLOAT/STR: code of the first
         element; ordered, not exact.

"Exact" codes let the device fold values without the host ever seeing the groups; otherwise the
device still does the partition+sort and the host walks the (tiny) runs of equal codes.  This
replaces Splitter.partition (base.py:6-8) and the key comparisons of list.sort / heapq.merge
(dataset.py:162-164, 571-579).
"""
import hashlib
import pickle
import struct

import numpy as np

INT, FLOAT, STR, BYTES, HASH, TUPLE = "int", "float", "str", "bytes", "hash", "tuple"
_BIAS = 1 << 63


class Codec(object):
    __slots__ = ("kind", "exact", "ordered", "decodable")

    def __init__(self, kind, decodable=False):
        self.kind = kind
        self.exact = kind in (INT, FLOAT)      # equal codes <=> equal keys
        self.ordered = kind != HASH
        # the key OBJECTS can be rebuilt from the codes: only when every key has exactly the type the
        # decoder produces (True == 1 == 1.0 and -0.0 == 0.0 are equal keys with one code, but a rebuilt key
        # must be the object the reference would show, so such inputs keep their original key objects)
        self.decodable = self.exact and decodable

    def __repr__(self):
        return "Codec(%s)" % self.kind


def _classify(keys):
    """Smallest codec that represents every key."""
    kinds = set()
    for k in keys:
        t = type(k)
        if t is int or t is bool:
            kinds.add(INT if -_BIAS <= k < _BIAS else HASH)
        elif t is float:
            kinds.add(FLOAT)
        elif t is str:
            kinds.add(STR)
        elif t is bytes:
            kinds.add(BYTES)
        elif isinstance(k, (np.integer,)):
            kinds.add(INT)
        elif isinstance(k, (np.floating,)):
            kinds.add(FLOAT)
        else:
            kinds.add(HASH)
        if HASH in kinds:
            return HASH
    if not kinds:
        return INT
    if len(kinds) == 1:
        return kinds.pop()
    if kinds == {INT, FLOAT}:
        if all(type(k) is float or abs(int(k)) <= (1 << 53) for k in keys):
            return FLOAT
    return HASH


def _float_codes(arr):
    bits = arr.astype(np.float64).view(np.uint64)
    neg = (bits >> np.uint64(63)).astype(bool)
    return np.where(neg, ~bits, bits | np.uint64(_BIAS))


def _prefix_codes(blobs):
    out = np.empty(len(blobs), dtype=np.uint64)
    for i, b in enumerate(blobs):
        out[i] = int.from_bytes(b[:8].ljust(8, b"\0"), "big")
    return out


def _canon(k):
    """Canonical form so that keys comparing equal hash alike (1 == 1.0 == True)."""
    if isinstance(k, (bool, int, np.integer)):
        return ("n", int(k))
    if isinstance(k, (float, np.floating)):
        f = float(k)
        return ("n", int(f)) if f == int(f) else ("f", f)
    if isinstance(k, tuple):
        return ("t",) + tuple(_canon(x) for x in k)
    if isinstance(k, frozenset):
        return ("s",) + tuple(sorted(repr(_canon(x)) for x in k))
    return k


def _hash_code(k):
    try:
        blob = pickle.dumps(_canon(k), protocol=4)
    except Exception:
        blob = repr(k).encode("utf-8", "replace")
    return struct.unpack("<Q", hashlib.blake2b(blob, digest_size=8).digest())[0]


def encode(keys, need_order=False, force=None):
    """keys: list -> (numpy uint64 codes, Codec)."""
    kind = force or _classify(keys)
    if kind == HASH and need_order:
        kind = _tuple_kind(keys)
    if kind == INT:
        arr = np.fromiter((int(k) for k in keys), dtype=np.int64, count=len(keys))
        return (arr.view(np.uint64) ^ np.uint64(_BIAS)), Codec(INT, all(type(k) is int for k in keys))
    if kind == FLOAT:
        arr = np.fromiter((float(k) for k in keys), dtype=np.float64, count=len(keys))
        zeros = arr == 0
        faithful = all(type(k) is float for k in keys) and not bool(np.signbit(arr[zeros]).any())
        arr = arr + 0.0   # -0.0 == 0.0: one key, one code
        return _float_codes(arr), Codec(FLOAT, faithful)
    if kind == STR:
        return _prefix_codes([k.encode("utf-8") for k in keys]), Codec(STR)
    if kind == BYTES:
        return _prefix_codes(keys), Codec(BYTES)
    if kind == TUPLE:
        firsts = [k[0] for k in keys]
        codes, _c = encode(firsts, need_order=True)
        return codes, Codec(TUPLE)
    codes = np.fromiter((_hash_code(k) for k in keys), dtype=np.uint64, count=len(keys))
    return codes, Codec(HASH)


def _tuple_kind(keys):
    if keys and all(isinstance(k, tuple) and len(k) > 0 for k in keys):
        sub = _classify([k[0] for k in keys])
        if sub != HASH:
            return TUPLE
        if _tuple_kind([k[0] for k in keys]) == TUPLE:
            return TUPLE
    raise TypeError("sort keys must be ints, floats, strings, bytes or tuples of them "
                    "(the device sort needs an order-preserving 64-bit prefix)")


def decode_exact(codes, codec):
    """Inverse of encode for exact codecs: numpy codes -> list of Python keys."""
    if codec.kind == INT:
        return (codes ^ np.uint64(_BIAS)).view(np.int64).tolist()
    if codec.kind == FLOAT:
        neg = (codes >> np.uint64(63)) == 0
        bits = np.where(neg, ~codes, codes & np.uint64(_BIAS - 1))
        return bits.view(np.float64).tolist()
    raise ValueError("codec %s is not exact" % codec.kind)


def refine_runs(keys, values, codes, codec):
    """Records already ordered by code; make equal KEYS adjacent (and ordered, when the codec is)
    inside every run of equal codes. Returns (keys, values) lists. Cheap: runs are tiny."""
    if codec.exact or len(keys) < 2:
        return keys, values
    same = codes[1:] == codes[:-1]
    if not same.any():
        return keys, values
    # run boundaries
    starts = np.flatnonzero(np.concatenate(([True], ~same)))
    ends = np.concatenate((starts[1:], [len(keys)]))
    keys = list(keys)
    values = list(values)
    for s, e in zip(starts.tolist(), ends.tolist()):
        if e - s < 2:
            continue
        idx = list(range(s, e))
        if codec.ordered:
            idx.sort(key=lambda i: keys[i])  # stable: ties keep device (input) order
        else:
            first = {}
            order = []
            for i in idx:
                ck = _canon(keys[i])
                try:
                    slot = first.setdefault(ck, len(first))
                except TypeError:
                    slot = first.setdefault(repr(ck), len(first))
                order.append((slot, i))
            order.sort()
            idx = [i for _s, i in order]
        ks = [keys[i] for i in idx]
        vs = [values[i] for i in idx]
        keys[s:e] = ks
        values[s:e] = vs
    return keys, values
