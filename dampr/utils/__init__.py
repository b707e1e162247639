from dampr_b200.utils import filter_by_count  # noqa: F401
