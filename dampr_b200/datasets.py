"""Datasets: what stages read and produce.

`Dataset` / `Chunker` keep the reference's public contract for custom inputs
(dampr/dataset.py:420-442: read() yields (key, value), delete(), chunks()).  The run files of the
reference (PickledDataset / MemGZipDataset, gzip'd pickles under /tmp) do not exist here: stage
outputs are `RecordsDataset` (host objects in device-sorted order) or `ColumnDataset` (numpy /
device columns materialised lazily).
"""
import gzip
import io
import itertools
import os

import numpy as np


class Chunker(object):
    def chunks(self):
        raise NotImplementedError()


class Dataset(Chunker):
    def read(self):
        raise NotImplementedError()

    def grouped_read(self):
        """(key, iterator of values) for runs of equal keys (dataset.py:429-433)."""
        for key, grp in itertools.groupby(self.read(), key=lambda kv: kv[0]):
            yield key, (kv[1] for kv in list(grp))

    def delete(self):
        pass

    def __iter__(self):
        return self.read()

    def chunks(self):
        yield self


class EmptyDataset(Dataset):
    def read(self):
        return iter(())


class MemoryDataset(Dataset):
    def __init__(self, kvs, partitions=13):
        self.kvs = kvs
        self.partitions = partitions

    def read(self):
        return iter(self.kvs)

    def chunks(self):
        n = len(self.kvs)
        if self.partitions <= 1 or n == 0:
            yield self
            return
        step = -(-n // self.partitions)
        for s in range(0, n, step):
            yield MemoryDataset(self.kvs[s:s + step], 1)


class StreamDataset(Dataset):
    def __init__(self, it):
        self.it = it

    def read(self):
        return self.it


class CatDataset(Dataset):
    def __init__(self, datasets):
        self.datasets = list(datasets)

    def read(self):
        for d in self.datasets:
            for kv in d.read():
                yield kv

    def delete(self):
        for d in self.datasets:
            d.delete()

    def chunks(self):
        return iter(self.datasets)


class RecordsDataset(Dataset):
    """Stage output on the host path: parallel lists of keys and values, already in the order the
    device sort produced (key order for orderable keys), equal keys adjacent."""

    def __init__(self, keys, values, codes=None, codec=None, replicated=False):
        self.keys = keys
        self.values = values
        self.codes = codes    # numpy uint64 per record (device sort key) or None
        self.codec = codec    # keycodec.Codec used for `codes`
        # under torch.distributed: True when every rank holds this same (global) result, e.g. a line count
        # summed over the ranks; False = this rank's shard of an owner-partitioned result
        self.replicated = replicated

    def __len__(self):
        return len(self.keys)

    def read(self):
        return zip(self.keys, self.values)

    def delete(self):
        self.keys, self.values, self.codes = [], [], None

    def chunks(self):
        yield self


class ColumnDataset(Dataset):
    """Columnar stage output of a lowered (device) stage: numpy key/value columns plus optional
    decoded string keys.  read() materialises Python tuples lazily, in row order."""

    def __init__(self, key_col, val_col, record_fn=None, n=None):
        self.key_col = key_col
        self.val_col = val_col
        self.record_fn = record_fn  # (i) -> (k, v) override
        self.n = len(key_col) if n is None else n

    def __len__(self):
        return self.n

    def read(self):
        if self.record_fn is not None:
            f = self.record_fn
            return (f(i) for i in range(self.n))
        kc = self.key_col.tolist() if isinstance(self.key_col, np.ndarray) else self.key_col
        vc = self.val_col.tolist() if isinstance(self.val_col, np.ndarray) else self.val_col
        return zip(kc, vc)

    def delete(self):
        self.key_col = self.val_col = ()
        self.n = 0


# ---- text inputs ---------------------------------------------------------------------------------
def _iter_lines(buf, base, universal=True):
    """(byte offset, str) for every line of `buf` (bytes). '\\n', '\\r\\n' and lone '\\r' end lines,
    a trailing unterminated piece is a line, like Python's universal-newline text mode. universal=False:
    binary-mode iteration (only '\\n' ends a line, a '\\r' stays in it) as the reference reads .gz files."""
    pos = 0
    n = len(buf)
    find = buf.find
    has_cr = universal and find(b"\r") >= 0
    while pos < n:
        if has_cr:
            i_n = find(b"\n", pos)
            i_r = find(b"\r", pos)
            if i_n < 0 and i_r < 0:
                end, nxt = n, n
            elif i_r >= 0 and (i_n < 0 or i_r < i_n):
                end = i_r
                nxt = i_r + 2 if (i_r + 1 < n and buf[i_r + 1:i_r + 2] == b"\n") else i_r + 1
            else:
                end, nxt = i_n, i_n + 1
        else:
            i_n = find(b"\n", pos)
            if i_n < 0:
                end, nxt = n, n
            else:
                end, nxt = i_n, i_n + 1
        yield base + pos, buf[pos:end].decode("utf-8")
        pos = nxt


class TextLineDataset(Dataset):
    """Byte range [start, end) of a text file. Owns exactly the lines whose first byte lies in the
    range (every line of the file belongs to exactly one chunk: the intended semantics of
    dataset.py:458-476, without its float-seam and long-line duplicates). Yields
    (byte offset of the line, line without its terminator)."""

    def __init__(self, path, start=0, end=None):
        self.path = path
        self.start = int(start)
        self.end = None if end is None else int(end)

    def read(self):
        size = os.path.getsize(self.path)
        start = min(self.start, size)
        end = size if self.end is None else min(self.end, size)
        if start >= end:
            return
        with open(self.path, "rb") as f:
            # first owned line start: `start` itself if the previous byte ends a line
            if start > 0:
                f.seek(start - 1)
                prev = f.read(1)
            else:
                prev = b"\n"
            f.seek(start)
            block = f.read(end - start)
            # the last owned line may continue past `end`
            tail = b""
            if end < size and not block.endswith((b"\n", b"\r")):
                while True:
                    more = f.read(1 << 16)
                    if not more:
                        break
                    cut = -1
                    for i, ch in enumerate(more):
                        if ch == 10 or ch == 13:
                            cut = i
                            break
                    if cut >= 0:
                        tail += more[:cut + 1]
                        break
                    tail += more
            elif end < size and block.endswith(b"\r"):
                nxt = f.read(1)  # "\r\n" split by the seam: the '\n' belongs to this terminator
                if nxt == b"\n":
                    tail = nxt
            first = 0
            if prev not in (b"\n", b"\r"):
                # skip the line that started in an earlier chunk
                i = 0
                n = len(block)
                while i < n and block[i] != 10 and block[i] != 13:
                    i += 1
                if i >= n:
                    return  # the whole chunk lies inside one line owned elsewhere
                first = i + 2 if (block[i] == 13 and block[i + 1:i + 2] == b"\n") else i + 1
            elif prev == b"\r" and block[:1] == b"\n":
                first = 1  # second half of a "\r\n" owned by the previous chunk
            body = block[first:]
            if not body:
                return
            # lines that start inside [start+first, end); the tail completes the last of them
            data = body + tail
            limit = end
            for off, line in _iter_lines(data, start + first):
                if off >= limit:
                    break
                yield off, line

    def __str__(self):
        return "Text[path=%s,start=%s,end=%s]" % (self.path, self.start, self.end)


class GzipLineDataset(Dataset):
    def __init__(self, path):
        self.path = path

    def read(self):
        # binary mode like the reference (dataset.py:488-493): split on '\n' only, rstrip(os.linesep) strips
        # just the '\n', so the '\r' of a CRLF file stays in the line
        with gzip.open(self.path, "rb") as f:
            data = f.read()
        return _iter_lines(data, 0, universal=False)

    def __str__(self):
        return "GzipFile[path=%s]" % self.path


class KVFileDataset(Dataset):
    """Binary file slice of little-endian 16-byte (u64 key, i64 value) records.
    read() yields (record index, (key, value)); columns() feeds the device path without touching
    Python objects."""

    def __init__(self, path, start_rec=0, end_rec=None, signed_keys=False):
        self.path = path
        self.start_rec = int(start_rec)
        total = os.path.getsize(path) // 16
        self.end_rec = total if end_rec is None else min(int(end_rec), total)
        self.signed_keys = signed_keys

    def records(self):
        n = max(0, self.end_rec - self.start_rec)
        return np.fromfile(self.path, dtype=np.uint64, count=2 * n, offset=16 * self.start_rec).reshape(n, 2)

    def columns(self):
        r = self.records()
        return r[:, 0].copy(), r[:, 1].copy().view(np.int64)

    def read(self):
        k, v = self.columns()
        if self.signed_keys:
            k = k.view(np.int64)
        for i, (a, b) in enumerate(zip(k.tolist(), v.tolist())):
            yield self.start_rec + i, (a, b)


class ArrayKVDataset(Dataset):
    """In-memory (keys, values) numpy columns presented as records (i, (key, value))."""

    def __init__(self, keys, vals, base=0):
        self.keys = np.ascontiguousarray(keys)
        self.vals = np.ascontiguousarray(vals)
        self.base = base

    def columns(self):
        return self.keys, self.vals

    def read(self):
        for i, (a, b) in enumerate(zip(self.keys.tolist(), self.vals.tolist())):
            yield self.base + i, (a, b)
