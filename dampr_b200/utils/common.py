"""DSL-level helpers (reference: dampr/utils/common.py:2-15)."""


def filter_by_count(pipe, key_func, filter_func):
    """Keep the items whose key occurs a number of times accepted by filter_func: count the keys,
    filter the counts, then join the surviving keys back onto the items (reduce-side join)."""
    kept = pipe.map(key_func).count().filter(lambda kc: filter_func(kc[1]))
    return kept.group_by(lambda kc: kc[0], lambda kc: kc[1]) \
        .join(pipe.group_by(key_func)) \
        .reduce(lambda _counts, items: items, many=True) \
        .map(lambda kv: kv[1])
