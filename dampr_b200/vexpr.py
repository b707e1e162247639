"""Column-at-a-time evaluation of lowered user lambdas over a frame.

The map stages that follow an aggregation (examples/word-stats.py:24-37: `fold_by(len(tc[0]), tc[1])`,
`.map(lambda wl: wl[0] * wl[1])`, `mean`) see one record per distinct key; in the reference they are
per-record Python like everything else (Map.stream, base.py:30-33).  Here the expression tree of the
lambda (lowering.analyze) is evaluated over whole numpy columns.  The result must be what CPython would
have produced for every row, so every operation is guarded:

  ints      int64 arithmetic only while the operands prove that no intermediate can leave +-2^62
            (Python ints do not wrap); true division only below 2^53 (float(a) / float(b) is then the
            correctly rounded quotient CPython computes); a zero divisor is left to the host path, which
            raises ZeroDivisionError like the reference.
  floats    IEEE double in both worlds: + - * / neg and comparisons are bit-identical.
  anything else (unknown calls, mixed-type min/max, strings other than len()) raises NotVec and the
  stage runs as a host map.
"""
import numpy as np

from . import lowering

E = lowering.E
LIM = 1 << 62


class NotVec(Exception):
    pass


class Const(object):
    """A value that is the same for every row."""
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v


class Tup(object):
    """A tuple-valued column (one entry per component)."""
    __slots__ = ("items",)

    def __init__(self, items):
        self.items = list(items)

    def __len__(self):
        return len(self.items[0]) if self.items else 0


def _is_num(c):
    return isinstance(c, np.ndarray) and c.dtype in (np.int64, np.float64)


def _as_array(c, n):
    from .plan import DictCol
    if isinstance(c, DictCol):
        if isinstance(c.uniq, np.ndarray):
            return c.uniq[c.inv]
        return c.materialize()
    return c


def _absmax(a):
    if isinstance(a, Const):
        return abs(a.v)
    if len(a) == 0:
        return 0
    lo, hi = int(a.min()), int(a.max())
    return max(abs(lo), abs(hi))


def _kind(a):
    if isinstance(a, Const):
        if type(a.v) is int:
            return "i"
        if type(a.v) is float:
            return "f"
        raise NotVec("constant of type %s" % type(a.v).__name__)
    if _is_num(a):
        return "i" if a.dtype == np.int64 else "f"
    raise NotVec("non-numeric operand")


def _val(a):
    return a.v if isinstance(a, Const) else a


def _to_float(a):
    """float(x) for an int operand (round-to-nearest in CPython and in numpy alike)."""
    if isinstance(a, Const):
        return Const(float(a.v))
    return a.astype(np.float64)


def _binop(op, l, r):
    kl, kr = _kind(l), _kind(r)
    if isinstance(l, Const) and isinstance(r, Const):
        raise NotVec("constant expression")  # let CPython fold it on the host path
    if op in ("+", "-", "*"):
        if kl == "i" and kr == "i":
            ml, mr = _absmax(l), _absmax(r)
            if (op == "*" and ml * mr >= LIM) or (op != "*" and ml + mr >= LIM):
                raise NotVec("int64 overflow possible")
        else:
            if kl == "i":
                if _absmax(l) >= (1 << 63):
                    raise NotVec("int too large for float conversion")
                l = _to_float(l)
            if kr == "i":
                if _absmax(r) >= (1 << 63):
                    raise NotVec("int too large for float conversion")
                r = _to_float(r)
        a, b = _val(l), _val(r)
        with np.errstate(all="ignore"):
            return a + b if op == "+" else (a - b if op == "-" else a * b)
    if op == "/":
        if kl == "i":
            if _absmax(l) >= (1 << 53):
                raise NotVec("int/ beyond 2^53")
            l = _to_float(l)
        if kr == "i":
            if _absmax(r) >= (1 << 53):
                raise NotVec("int/ beyond 2^53")
            r = _to_float(r)
        b = _val(r)
        if (b == 0) if isinstance(r, Const) else bool((b == 0).any()):
            raise NotVec("division by zero")  # the host path raises ZeroDivisionError like the reference
        with np.errstate(all="ignore"):
            return _val(l) / b
    if op in ("//", "%"):
        if kl != "i" or kr != "i":
            raise NotVec("float floor division")
        b = _val(r)
        if (b == 0) if isinstance(r, Const) else bool((b == 0).any()):
            raise NotVec("division by zero")
        if _absmax(l) >= LIM or _absmax(r) >= LIM:
            raise NotVec("int64 overflow possible")
        a = _val(l)
        return np.floor_divide(a, b) if op == "//" else np.mod(a, b)  # floor semantics, like CPython
    raise NotVec("operator %s" % op)


_CMP = {"<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal, "==": np.equal, "!=": np.not_equal}


def _cmp(op, l, r):
    if op not in _CMP:
        raise NotVec("comparison %s" % op)
    kl, kr = _kind(l), _kind(r)
    if isinstance(l, Const) and isinstance(r, Const):
        raise NotVec("constant expression")
    if kl != kr:
        # CPython compares int with float exactly; below 2^53 the conversion is exact too
        if (kl == "i" and _absmax(l) >= (1 << 53)) or (kr == "i" and _absmax(r) >= (1 << 53)):
            raise NotVec("int/float comparison beyond 2^53")
        if kl == "i":
            l = _to_float(l)
        else:
            r = _to_float(r)
    return _CMP[op](_val(l), _val(r))


def _strlen(col):
    """len() of every string of a column."""
    from .plan import DictCol
    if isinstance(col, DictCol):
        u = _strlen(col.uniq)
        return u[col.inv]
    if isinstance(col, np.ndarray) and col.dtype.kind == "S":
        return np.char.str_len(col).astype(np.int64)
    if isinstance(col, list):
        if not all(type(x) is str for x in col):
            raise NotVec("len() of a non-string")
        return np.fromiter((len(x) for x in col), dtype=np.int64, count=len(col))
    raise NotVec("len() of a non-string column")


def evaluate(e, cols, scalar, n):
    """Value of expression `e` (argument 0 = the record) over a frame's columns: a numpy array, a
    string column passed through unchanged, a Const or a Tup. Raises NotVec."""

    def ev(x):
        if not isinstance(x, E):
            raise NotVec("unsupported node")
        op = x.op
        if op == "arg":
            if x.a != 0:
                raise NotVec("second argument")
            return cols[0] if scalar else Tup(cols)
        if op == "const":
            if type(x.a) in (int, float):
                return Const(x.a)
            raise NotVec("constant of type %s" % type(x.a).__name__)
        if op == "sub":
            base = ev(x.a)
            ok, i = lowering._const(x.b)
            if not ok or type(i) is not int or not isinstance(base, Tup) or not (0 <= i < len(base.items)):
                raise NotVec("subscript")
            return base.items[i]
        if op == "tuple":
            return Tup([ev(y) for y in x.a])
        if op == "neg":
            v = ev(x.a)
            k = _kind(v)
            if isinstance(v, Const):
                return Const(-v.v)
            if k == "i" and _absmax(v) >= LIM:
                raise NotVec("int64 overflow possible")
            return -v
        if op == "bin":
            return _binop(x.a, _num(ev(x.b)), _num(ev(x.c)))
        if op == "cmp":
            return _cmp(x.a, _num(ev(x.b)), _num(ev(x.c)))
        if op == "call" and isinstance(x.a, E) and x.a.op == "obj" and len(x.b) == 1:
            f = x.a.a
            if f is len:
                return _strlen(ev(x.b[0]))
            if f is float:
                v = _num(ev(x.b[0]))
                if _kind(v) == "i":
                    if _absmax(v) >= (1 << 63):
                        raise NotVec("int too large for float conversion")
                    return _to_float(v)
                return v
            if f is abs:
                v = _num(ev(x.b[0]))
                if isinstance(v, Const):
                    return Const(abs(v.v))
                if _kind(v) == "i" and _absmax(v) >= LIM:
                    raise NotVec("int64 overflow possible")
                return np.abs(v)
        raise NotVec("unsupported expression %r" % (x,))

    def _num(v):
        v = _as_array(v, n) if not isinstance(v, (Const, Tup)) else v
        if isinstance(v, Tup):
            raise NotVec("tuple operand")
        if isinstance(v, np.ndarray) and v.dtype == np.uint64:
            if len(v) and int(v.max()) >= (1 << 63):
                raise NotVec("uint64 beyond int64")
            v = v.astype(np.int64)
        if isinstance(v, np.ndarray) and v.dtype == np.bool_:
            raise NotVec("bool operand")
        return v

    return ev(e)


def broadcast(v, n):
    """Materialise an evaluate() result as a frame column (Const -> full array)."""
    if isinstance(v, Const):
        return np.full(n, v.v, dtype=np.int64 if type(v.v) is int else np.float64)
    return v
