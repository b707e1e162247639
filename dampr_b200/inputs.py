"""Input chunkers (reference: dampr/inputs.py:14-97)."""
import glob
import os

import numpy as np

from .datasets import (Chunker, Dataset, TextLineDataset, GzipLineDataset, MemoryDataset, KVFileDataset,
                       ArrayKVDataset)


def read_paths(paths, follow_links):
    """Files named by a glob / directory / list of them, dot-files skipped (inputs.py:14-30)."""
    if not isinstance(paths, (list, tuple)):
        paths = [paths]
    for pattern in paths:
        for path in glob.glob(pattern):
            if os.path.isfile(path):
                cands = [path]
            else:
                cands = (os.path.join(root, f)
                         for root, _dirs, files in os.walk(path, followlinks=follow_links) for f in files)
            for p in cands:
                if not os.path.basename(p).startswith("."):
                    yield p


class TextInput(Chunker):
    """One file split into byte ranges of chunk_size; a .gz file is one unsplittable chunk."""

    def __init__(self, path, chunk_size=64 * 1024 ** 2):
        self.path = path
        self.chunk_size = chunk_size

    def chunks(self):
        if self.path.endswith(".gz"):
            yield GzipLineDataset(self.path)
            return
        size = os.path.getsize(self.path)
        step = max(1, int(self.chunk_size))  # the tf-idf script passes a float (tf-idf-dampr.py:9-10)
        for off in range(0, size, step):
            yield TextLineDataset(self.path, off, min(size, off + step))


class PathInput(Chunker):
    def __init__(self, path, chunk_size=64 * 1024 ** 2, follow_links=True):
        self.path = path
        self.chunk_size = chunk_size
        self.follow_links = follow_links

    def files(self):
        return list(read_paths(self.path, self.follow_links))

    def chunks(self):
        for p in self.files():
            for c in TextInput(p, self.chunk_size).chunks():
                yield c


class MemoryInput(Chunker):
    def __init__(self, items, partitions=50):
        self.items = items
        self.partitions = min(len(items), partitions)

    def chunks(self):
        if self.partitions == 0:
            yield MemoryDataset(self.items)
            return
        step = max(1, len(self.items) // self.partitions)
        for s in range(0, len(self.items), step):
            yield MemoryDataset(self.items[s:s + step])


class UrlsInput(Chunker):
    def __init__(self, urls, skip_on_error=True):
        self.urls = urls
        self.soe = skip_on_error

    def chunks(self):
        for u in self.urls:
            yield UrlDataset(u, self.soe)


class UrlDataset(Dataset):
    def __init__(self, path, skip_on_error=True):
        self.path = path
        self.soe = skip_on_error

    def read(self):
        from urllib.request import urlopen
        from urllib.error import HTTPError
        try:
            with urlopen(self.path) as h:
                for i, line in enumerate(h):
                    yield i, line.decode("utf-8")
        except HTTPError:
            if not self.soe:
                raise


class KVInput(Chunker):
    """Binary 16-byte record file (u64 key, i64 value) cut into record-aligned chunks. The reference
    would be fed the same file through a custom Chunker of Datasets; here the chunks
    also expose numpy columns so lowered stages never build Python tuples."""

    def __init__(self, path, chunk_records=4 << 20):
        self.path = path
        self.chunk_records = int(chunk_records)

    def n_records(self):
        return os.path.getsize(self.path) // 16

    def chunks(self):
        n = self.n_records()
        for s in range(0, n, self.chunk_records):
            yield KVFileDataset(self.path, s, min(n, s + self.chunk_records))

    def columns(self):
        r = np.fromfile(self.path, dtype=np.uint64).reshape(-1, 2)
        return r[:, 0].copy(), r[:, 1].copy().view(np.int64)


class ArrayKVInput(Chunker):
    """(keys, values) numpy arrays as an input of (i, (key, value)) records."""

    def __init__(self, keys, vals, chunk_records=4 << 20):
        self.keys = np.ascontiguousarray(keys)
        self.vals = np.ascontiguousarray(vals)
        self.chunk_records = int(chunk_records)

    def chunks(self):
        n = len(self.keys)
        for s in range(0, n, self.chunk_records):
            e = min(n, s + self.chunk_records)
            yield ArrayKVDataset(self.keys[s:e], self.vals[s:e], s)

    def columns(self):
        return self.keys, self.vals
