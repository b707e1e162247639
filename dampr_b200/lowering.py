"""Lowering of user lambdas: CPython bytecode -> small expression trees -> closed set of idioms.

The device cannot run Python, and a map left in CPython ingests ~10 MB/s/core, so at
graph-run time the planner inspects the user functions of a stage.  A function is lowered only if
its whole body is a straight-line expression over its arguments built from the whitelisted forms
below — anything else (branches, loops, unknown calls, unknown opcodes) returns None and the stage
runs as a host map that feeds the device shuffle.  The analysis is sound by construction: it never
guesses from observed outputs.

Idioms that matter for the five BASELINE configs (SURVEY "Hard parts"):
  x.split()                                   examples/wc.py:12
  set(RX.split(x.lower())), RX = [^\\w]+       benchmarks/tf-idf-dampr.py:12-14
  lambda x: x / lambda x: 1 / x[i] / -x[i]    key and value projections
  lambda x, y: x + y, operator.add, min, max  associative folds (dampr.py:661-708)
  lambda k, it: sum(it) / len(list(it))       group reducers
"""
import builtins
import dis
import operator
import re
import types


class E(object):
    """Expression node."""
    __slots__ = ("op", "a", "b", "c")

    def __init__(self, op, a=None, b=None, c=None):
        self.op, self.a, self.b, self.c = op, a, b, c

    def __repr__(self):
        parts = [repr(x) for x in (self.a, self.b, self.c) if x is not None]
        return "%s(%s)" % (self.op, ", ".join(parts))

    def __eq__(self, other):
        return isinstance(other, E) and (self.op, self.a, self.b, self.c) == (other.op, other.a, other.b, other.c)

    def __hash__(self):
        return hash((self.op, repr(self.a), repr(self.b), repr(self.c)))


# node kinds: arg(i) const(v) obj(pyobject) attr(e, name) call(f, [args]) bin(op, l, r) neg(e)
#             sub(e, idx) tuple([items]) cmp(op, l, r)
_NULL = object()


_INSTR_CACHE = {}


def _instructions(code):
    """Decoded bytecode of a code object (cached: graphs are rebuilt with fresh lambdas on every run, but
    their code objects are the same, and decoding is the expensive part of the analysis)."""
    ins = _INSTR_CACHE.get(code)
    if ins is None:
        if len(_INSTR_CACHE) > 4096:
            _INSTR_CACHE.clear()
        ins = _INSTR_CACHE[code] = list(dis.get_instructions(code))
    return ins


def analyze(fn):
    """Expression tree of a plain Python function / lambda, or None if it is not a straight-line
    expression over whitelisted opcodes."""
    if not isinstance(fn, types.FunctionType):
        return None
    code = fn.__code__
    if code.co_flags & (0x20 | 0x80 | 0x100 | 0x200):  # generator / coroutine variants
        return None
    if code.co_kwonlyargcount or (code.co_flags & 0x0C):  # *args / **kwargs
        return None
    nargs = code.co_argcount
    argnames = code.co_varnames[:nargs]
    stack = []
    try:
        for ins in _instructions(code):
            name = ins.opname
            if name in ("RESUME", "NOP", "COPY_FREE_VARS", "MAKE_CELL", "PRECALL", "CACHE"):
                continue
            if name in ("LOAD_FAST", "LOAD_FAST_CHECK"):
                if ins.argval not in argnames:
                    return None
                stack.append(E("arg", argnames.index(ins.argval)))
            elif name == "LOAD_CONST":
                stack.append(E("const", ins.argval))
            elif name == "RETURN_CONST":
                return E("const", ins.argval)
            elif name == "LOAD_GLOBAL":
                if ins.arg & 1:
                    stack.append(_NULL)
                gname = ins.argval
                if gname in fn.__globals__:
                    stack.append(E("obj", fn.__globals__[gname]))
                elif hasattr(builtins, gname):
                    stack.append(E("obj", getattr(builtins, gname)))
                else:
                    return None
            elif name == "LOAD_DEREF":
                idx = (code.co_cellvars + code.co_freevars).index(ins.argval) - len(code.co_cellvars)
                if idx < 0 or fn.__closure__ is None:
                    return None
                stack.append(E("obj", fn.__closure__[idx].cell_contents))
            elif name == "PUSH_NULL":
                stack.append(_NULL)
            elif name == "LOAD_ATTR":
                obj = stack.pop()
                if ins.arg & 1:
                    stack.append(_NULL)
                stack.append(E("attr", obj, ins.argval))
            elif name == "LOAD_METHOD":
                obj = stack.pop()
                stack.append(_NULL)
                stack.append(E("attr", obj, ins.argval))
            elif name == "CALL":
                n = ins.arg
                args = stack[len(stack) - n:] if n else []
                del stack[len(stack) - n:]
                f = stack.pop()
                if stack and stack[-1] is _NULL:
                    stack.pop()
                if any(a is _NULL for a in args) or f is _NULL:
                    return None
                stack.append(E("call", f, tuple(args)))
            elif name == "BINARY_OP":
                r = stack.pop()
                l = stack.pop()
                stack.append(E("bin", ins.argrepr.rstrip("="), l, r))
            elif name == "BINARY_SUBSCR":
                i = stack.pop()
                o = stack.pop()
                stack.append(E("sub", o, i))
            elif name == "UNARY_NEGATIVE":
                stack.append(E("neg", stack.pop()))
            elif name == "BUILD_TUPLE":
                n = ins.arg
                items = tuple(stack[len(stack) - n:]) if n else ()
                del stack[len(stack) - n:]
                stack.append(E("tuple", items))
            elif name == "COMPARE_OP":
                r = stack.pop()
                l = stack.pop()
                stack.append(E("cmp", ins.argval, l, r))
            elif name == "CONTAINS_OP":
                r = stack.pop()
                l = stack.pop()
                stack.append(E("cmp", "not in" if ins.arg else "in", l, r))
            elif name == "RETURN_VALUE":
                out = stack.pop()
                return None if out is _NULL else out
            else:
                return None
    except (IndexError, ValueError, AttributeError):
        return None
    return None


# ---- idiom matchers ----------------------------------------------------------------------------
def _is_arg(e, i=0):
    return isinstance(e, E) and e.op == "arg" and e.a == i


def _const(e):
    """(True, value) if e is a constant."""
    if isinstance(e, E) and e.op == "const":
        return True, e.a
    return False, None


def is_identity(fn):
    e = analyze(fn)
    return e is not None and _is_arg(e, 0) and fn.__code__.co_argcount == 1


def constant_value(fn):
    """(True, v) if fn ignores its argument and returns the constant v."""
    e = analyze(fn)
    if e is None:
        return False, None
    return _const(e)


_NONWORD_PATTERNS = (r"[^\w]+", r"\W+")


def _is_nonword_pattern(obj):
    return isinstance(obj, re.Pattern) and obj.pattern in _NONWORD_PATTERNS and \
        (obj.flags & ~re.UNICODE) == 0


def _is_lower_of_arg(e):
    return isinstance(e, E) and e.op == "call" and e.b == () and isinstance(e.a, E) and \
        e.a.op == "attr" and e.a.b == "lower" and _is_arg(e.a.a)


def _is_nonword_split(e):
    """RX.split(x.lower()) or re.split(PATTERN, x.lower())"""
    if not (isinstance(e, E) and e.op == "call"):
        return False
    f, args = e.a, e.b
    if isinstance(f, E) and f.op == "attr" and f.b == "split" and isinstance(f.a, E) and f.a.op == "obj":
        if _is_nonword_pattern(f.a.a) and len(args) == 1 and _is_lower_of_arg(args[0]):
            return True
        if f.a.a is re and len(args) == 2:
            ok, pat = _const(args[0])
            return ok and pat in _NONWORD_PATTERNS and _is_lower_of_arg(args[1])
    if isinstance(f, E) and f.op == "obj" and f.a is re.split and len(args) == 2:
        ok, pat = _const(args[0])
        return ok and pat in _NONWORD_PATTERNS and _is_lower_of_arg(args[1])
    return False


TOK_WS, TOK_NONWORD_LOWER_SET, TOK_NONWORD_LOWER = 0, 1, 2


def tokenizer_mode(fn):
    """Device tokeniser mode of a flat_map function over a text line, or None."""
    e = analyze(fn)
    if e is None or fn.__code__.co_argcount != 1:
        return None
    # x.split()
    if e.op == "call" and e.b == () and isinstance(e.a, E) and e.a.op == "attr" and e.a.b == "split" \
            and _is_arg(e.a.a):
        return TOK_WS
    if _is_nonword_split(e):
        return TOK_NONWORD_LOWER
    if e.op == "call" and isinstance(e.a, E) and e.a.op == "obj" and e.a.a in (set, frozenset) \
            and len(e.b) == 1 and _is_nonword_split(e.b[0]):
        return TOK_NONWORD_LOWER_SET
    return None


def field_index(fn):
    """i if fn is `lambda x: x[i]` with a constant non-negative int i, else None."""
    e = analyze(fn)
    return _field(e)


def _field(e):
    if isinstance(e, E) and e.op == "sub" and _is_arg(e.a):
        ok, i = _const(e.b)
        if ok and type(i) is int and i >= 0:
            return i
    return None


def projection(fn):
    """Numeric projection of one record: ('field', i, sign) for x[i] / -x[i], ('ident', None, sign)
    for x / -x, ('const', v, 1) for constants; None otherwise."""
    e = analyze(fn)
    if e is None or fn.__code__.co_argcount != 1:
        return None
    sign = 1
    if e.op == "neg":
        sign, e = -1, e.a
    i = _field(e)
    if i is not None:
        return ("field", i, sign)
    if _is_arg(e):
        return ("ident", None, sign)
    ok, v = _const(e)
    if ok and sign == 1:
        return ("const", v, 1)
    return None


ADD, MIN, MAX, FIRST, LAST = "add", "min", "max", "first", "last"


def binop_kind(fn):
    """Kind of an associative binary operator, or None."""
    if fn is operator.add:
        return ADD
    if fn is min:
        return MIN
    if fn is max:
        return MAX
    e = analyze(fn)
    if e is None or getattr(fn, "__code__", None) is None or fn.__code__.co_argcount != 2:
        return None
    if e.op == "bin" and e.a == "+" and ((_is_arg(e.b, 0) and _is_arg(e.c, 1)) or (_is_arg(e.b, 1) and _is_arg(e.c, 0))):
        return ADD
    if e.op == "call" and isinstance(e.a, E) and e.a.op == "obj" and len(e.b) == 2 and \
            {x.a for x in e.b if isinstance(x, E) and x.op == "arg"} == {0, 1}:
        if e.a.a is min:
            return MIN
        if e.a.a is max:
            return MAX
        if e.a.a is operator.add:
            return ADD
    if _is_arg(e, 0):
        return FIRST
    if _is_arg(e, 1):
        return LAST
    return None


def tuple_binop_kinds(fn):
    """[kind per component] of a component-wise binary operator over tuples, e.g. mean()'s
    `lambda x, y: (x[0] + y[0], x[1] + y[1])` (dampr.py:450-467) -> [ADD, ADD]; None otherwise."""
    e = analyze(fn)
    if e is None or getattr(fn, "__code__", None) is None or fn.__code__.co_argcount != 2 or e.op != "tuple":
        return None
    out = []
    for i, comp in enumerate(e.a):
        if not (isinstance(comp, E) and comp.op == "bin" and comp.a == "+"):
            return None
        sides = []
        for side in (comp.b, comp.c):
            if not (isinstance(side, E) and side.op == "sub" and isinstance(side.a, E) and side.a.op == "arg"):
                return None
            ok, j = _const(side.b)
            if not ok or j != i:
                return None
            sides.append(side.a.a)
        if sorted(sides) != [0, 1]:
            return None
        out.append(ADD)
    return out or None


SUM, COUNT = "sum", "count"


def group_reducer_kind(fn):
    """Kind of a group reducer f(key, values_iter): sum(it) / len(list(it)) / min(it) / max(it)."""
    e = analyze(fn)
    if e is None or fn.__code__.co_argcount != 2:
        return None
    if e.op == "call" and isinstance(e.a, E) and e.a.op == "obj" and len(e.b) == 1:
        f, a = e.a.a, e.b[0]
        if _is_arg(a, 1):
            if f is sum:
                return SUM
            if f is min:
                return MIN
            if f is max:
                return MAX
        if f is len and isinstance(a, E) and a.op == "call" and isinstance(a.a, E) and a.a.op == "obj" \
                and a.a.a in (list, tuple) and len(a.b) == 1 and _is_arg(a.b[0], 1):
            return COUNT
    return None


def _iter_fold_kind(e, argi):
    """kind of sum(arg) / min(arg) / max(arg) / len(list(arg)) over argument `argi`, else None"""
    if not (isinstance(e, E) and e.op == "call" and isinstance(e.a, E) and e.a.op == "obj" and len(e.b) == 1):
        return None
    f, a = e.a.a, e.b[0]
    if _is_arg(a, argi):
        return {sum: SUM, min: MIN, max: MAX}.get(f) if f in (sum, min, max) else None
    if f is len and isinstance(a, E) and a.op == "call" and isinstance(a.a, E) and a.a.op == "obj" \
            and a.a.a in (list, tuple) and len(a.b) == 1 and _is_arg(a.b[0], argi):
        return COUNT
    return None


PRODUCT = "product"


def join_aggregate_kind(fn):
    """Shape of a join aggregate f(left_values_iter, right_values_iter) (dampr.py:780-820):
      ("folds", kl, kr)   lambda l, r: (K(l), K'(r)) with K, K' in sum / min / max / len(list(.))
      ("product",)        lambda l, r: itertools.product(l, r)   (one output per pair: use with many=True)
    None for anything else (the join then walks Python lists on the host)."""
    import itertools
    e = analyze(fn)
    if e is None or getattr(fn, "__code__", None) is None or fn.__code__.co_argcount != 2:
        return None
    if e.op == "tuple" and len(e.a) == 2:
        kl, kr = _iter_fold_kind(e.a[0], 0), _iter_fold_kind(e.a[1], 1)
        if kl is not None and kr is not None:
            return ("folds", kl, kr)
        return None
    if e.op == "call" and isinstance(e.a, E) and len(e.b) == 2 and _is_arg(e.b[0], 0) and _is_arg(e.b[1], 1):
        f = e.a
        callee = f.a if f.op == "obj" else (
            getattr(f.a.a, f.b, None) if (f.op == "attr" and isinstance(f.a, E) and f.a.op == "obj" and
                                          isinstance(f.a.a, types.ModuleType)) else None)
        if callee is itertools.product:
            return (PRODUCT,)
    return None


def depends_on(e, argi, fields=None):
    """Which fields of argument `argi` an expression reads: set of ints, or None when the argument is
    used whole (or in a way the analysis does not follow)."""
    out = set()

    def walk(x):
        if not isinstance(x, E):
            if isinstance(x, tuple):
                return all(walk(y) for y in x)
            return True
        if x.op == "sub" and _is_arg(x.a, argi):
            ok, i = _const(x.b)
            if ok and type(i) is int:
                out.add(i)
                return True
            return False
        if x.op == "arg":
            return x.a != argi
        return all(walk(y) for y in (x.a, x.b, x.c) if y is not None)

    return out if walk(e) else None
