"""Runs routines of the REFERENCE's own Fortran source here, where no Fortran compiler exists (test infrastructure).

A small Fortran-90 subset -> Python transpiler: one SUBROUTINE of a reference file (read in place from /root/reference,
never copied) becomes a Python function that executes the same statements in the same order with real*4 semantics --
every REAL operation is one numpy float32 operation (round to nearest, no contraction: what gfortran -O2 emits for
x86-64 without -ffast-math), INTEGER arithmetic is exact with Fortran's truncating division, x**n is the multiply
sequence of __powisf2, transcendental intrinsics go to the host libm's float routines (cosf, atanf, powf, ...), which is
what gfortran links.  The result is the reference's routine run on inputs of our choosing: tests/test_reference_source.py,
test_reference_init.py and test_reference_bdy_host.py compare such runs with the oracle, with wrf_b200/shipped.py and with
wrf_b200/bdy.py bit for bit, and tests/golden/reference_source.json, feeder_columns.json, shipped_init.json keep digests of what
the reference produced for the machines that have no reference tree.

Subset (whatever the ~120 routines of the path use; anything else raises NotImplementedError when a routine is translated):
  declarations  REAL / INTEGER / LOGICAL / DOUBLE PRECISION / REAL(KIND=8), old style without `::`, PARAMETER attribute and statement,
                initialised (SAVEd) locals, OPTIONAL, CHARACTER, implicit typing where IMPLICIT NONE is missing, declarations after
                statement functions, TYPE(...) dummies and locals
  statements    assignment (elements, whole arrays, a(:,:,:) = scalar), statement functions, DO / DO WHILE / END DO (named or not,
                CYCLE / EXIT), block and one-line IF, SELECT CASE (values, ranges, DEFAULT), RETURN, `;`, WRITE / PRINT as no-ops,
                CALL with positional, keyword and OPTIONAL arguments, scalars copied back, arrays by reference
  association   an element as the start of a slab (moist(ims,kms,jms,im) -> a 3-D view), rank-changing dummies (2-D mu through
                (ims:ime, 1:1, jms:jme)), the VPOW vector call of frame/libmassv.F
  expressions   Fortran precedence and typing rules, mixed real*4 / real*8 / integer promotion, .EQV. / .NEQV., intrinsics
                (MAX, MIN, ABS, SIGN, MOD, DIM, INT, NINT, REAL, SNGL, DBLE, ICHAR, CHAR, TRIM, PRESENT, the libm set)
  program units internal FUNCTIONs, callees translated on first use (Library), Python hooks for what lies outside the subset (namelist
                getters, message routines, file readers), module PARAMETERs and module variables, derived-type objects with array
                components (grid%u_2(i,k,j), model_config_rec%...), a line range of a larger routine wrapped as a routine of its own
                (Library.add_fragment), declarations standing in for a build-time #include (Library.inject), cpp conditionals
One limitation to know about: the source is folded to lower case including character literals ('W' reads 'w'); callers pass
character arguments in lower case.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import os
import re

import numpy as np

f32 = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
for _n in ("cosf", "sinf", "atanf", "expf", "logf", "sqrtf", "tanhf", "acosf", "asinf", "tanf"):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
_libm.powf.restype = ctypes.c_float
_libm.powf.argtypes = [ctypes.c_float, ctypes.c_float]


def _m1(name):
    fn = getattr(_libm, name)
    return lambda x: f32(fn(float(f32(x))))


def _idiv(a, b):          # Fortran integer division truncates towards zero
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def _powi(x, n):          # libgcc __powisf2: the multiply sequence gfortran uses for real ** integer
    m = abs(int(n))
    T = np.float64 if isinstance(x, np.float64) else f32        # __powidf2 for a double precision base
    x = T(x)
    y = x if (m & 1) else T(1.0)
    m >>= 1
    while m:
        x = T(x * x)
        if m & 1:
            y = T(y * x)
        m >>= 1
    return T(T(1.0) / y) if n < 0 else y


def _sign(a, b):
    neg = (b < 0) or (isinstance(b, (np.floating, float)) and b == 0 and np.signbit(b))
    return -abs(a) if neg else abs(a)


def _fit(a, rank, shape_of):
    """the actual array seen through the dummy's declared shape when the ranks differ (a view: only extent-1 axes come or go)"""
    if a is None or not isinstance(a, np.ndarray) or a.ndim == rank:
        return a
    shape = shape_of()
    if [n for n in a.shape if n != 1] != [n for n in shape if n != 1]:
        raise NotImplementedError(f"sequence association of shape {a.shape} with {shape}")
    v = a.reshape(shape)
    assert np.shares_memory(v, a)
    return v


class FArr:
    """an array component of a derived-type object with its Fortran lower bounds (grid%u_2 -> (ims, kms, jms))"""

    def __init__(self, a, lo):
        self.a, self.lo = a, tuple(lo)

    def idx(self, subs):
        if len(subs) != len(self.lo):
            raise IndexError(f"rank {len(self.lo)} component indexed with {len(subs)} subscripts")
        return tuple(int(s) - l for s, l in zip(subs, self.lo))


def _oget(obj, comp, subs):
    if comp not in obj and hasattr(obj, "make"):
        obj.make(comp, None if subs is None else len(subs))
    v = obj[comp]
    if subs is None:
        return v.a if isinstance(v, FArr) else v
    return v.a[v.idx(subs)]


def _oset(obj, comp, subs, val):
    if comp not in obj and hasattr(obj, "make") and subs is not None:
        obj.make(comp, len(subs))
    if subs is None:
        if isinstance(obj.get(comp), FArr):
            obj[comp].a[...] = val
        else:
            obj[comp] = val
    else:
        v = obj[comp]
        v.a[v.idx(subs)] = val


class _Obj(dict):
    """a derived-type local the routine fills itself (config_flags of init_domain_rk, set by model_to_grid_config_rec)"""


def _dim(a, b):
    return (a - b) if a > b else (a - a)


RUNTIME = dict(_fit=_fit, _dim=_dim, _oget=_oget, _oset=_oset, _Obj=_Obj, f32=f32, f64=np.float64, _idiv=_idiv, _powi=_powi, _sign=_sign, _cos=_m1("cosf"), _sin=_m1("sinf"), _atan=_m1("atanf"),
               _exp=_m1("expf"), _log=_m1("logf"), _sqrt=_m1("sqrtf"), _tanh=_m1("tanhf"), _acos=_m1("acosf"), _asin=_m1("asinf"), _tan=_m1("tanf"),
               _powf=lambda a, b: f32(_libm.powf(float(f32(a)), float(f32(b)))), np=np)


class _Loop(Exception):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# source preparation
def _cpp(lines, defines):
    out, stack = [], []          # stack of [taking, taken_any, parent_taking]

    def cond(expr):
        e = re.sub(r"defined\s*\(\s*(\w+)\s*\)", lambda m: "1" if m.group(1) in defines else "0", expr)
        e = re.sub(r"[A-Za-z_]\w*", lambda m: str(defines.get(m.group(0), 0)), e)
        e = e.replace("&&", " and ").replace("||", " or ").replace("!", " not ").replace(" not =", "!=")
        try:
            return bool(eval(e, {}, {}))
        except Exception:       # noqa: BLE001
            return False

    for ln in lines:
        s = ln.strip()
        if s.startswith("#"):
            d = s[1:].strip()
            parent = all(t[0] for t in stack)
            if d.startswith("ifdef"):
                v = d.split()[1] in defines
                stack.append([v, v])
            elif d.startswith("ifndef"):
                v = d.split()[1] not in defines
                stack.append([v, v])
            elif d.startswith("if"):
                v = cond(d[2:])
                stack.append([v, v])
            elif d.startswith("elif"):
                v = (not stack[-1][1]) and cond(d[4:])
                stack[-1][0] = v
                stack[-1][1] = stack[-1][1] or v
            elif d.startswith("else"):
                stack[-1][0] = not stack[-1][1]
                stack[-1][1] = True
            elif d.startswith("endif"):
                stack.pop()
            del parent
            continue
        if all(t[0] for t in stack):
            out.append(ln)
    return out


def _split_statements(text):
    """a ; b ; c -> [a, b, c], leaving the ; inside character literals alone"""
    out, cur, q = [], "", None
    for ch in text:
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == ";":
            out.append(cur)
            cur = ""
            continue
        cur += ch
    out.append(cur)
    return out


def _strip_comment(ln):
    q = None
    for n, ch in enumerate(ln):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return ln[:n]
    return ln


def statements(path, name, defines=None, text=None):
    """the statements of SUBROUTINE `name` in file `path` (or in `text`): comments dropped, continuations joined, lower case"""
    lines = _cpp((text if text is not None else open(path, errors="replace").read()).split("\n"), defines or {})
    lines = [_strip_comment(ln).rstrip() for ln in lines]
    out, cur, on = [], "", False
    start = re.compile(r"^\s*((real|integer|logical)\s+)?(subroutine|function)\s+%s\b" % re.escape(name), re.I)
    end = re.compile(r"^\s*end\s*(subroutine|function)(\s+%s\b|\s*$)" % re.escape(name), re.I)
    for ln in lines:
        if not ln.strip():
            continue
        if not on and not cur:
            if not start.match(ln):
                continue
            on = True
        text = ln.strip()
        if cur:
            text = text[1:].lstrip() if text.startswith("&") else text
            cur = cur + " " + text
        else:
            cur = text
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        for st in _split_statements(cur):
            if st.strip():
                out.append(st.strip().lower())
        cur = ""
        if end.match(out[-1]):
            return out
    raise KeyError(f"SUBROUTINE {name} not found in {path}")


# ---------------------------------------------------------------------------------------------------------------------
# expressions
_TOK = re.compile(r"\s*('[^']*'|\"[^\"]*\"|\d+\.(?!(?:lt|le|gt|ge|eq|ne|and|or|not|eqv|neqv)\.)\d*(?:[ed][+-]?\d+)?|\.\d+(?:[ed][+-]?\d+)?|\d+(?:[ed][+-]?\d+)?|\.(?:lt|le|gt|ge|eq|ne|and|or|not|true|false|eqv|neqv)\."
                  r"|[a-z_]\w*|\*\*|==|/=|<=|>=|[-+*/()<>,:%=])")
_REL = {".lt.": "<", ".le.": "<=", ".gt.": ">", ".ge.": ">=", ".eq.": "==", ".ne.": "!=", "<": "<", "<=": "<=", ">": ">", ">=": ">=", "==": "==",
        "/=": "!="}
_INTR1 = {"cos": "_cos", "sin": "_sin", "atan": "_atan", "exp": "_exp", "log": "_log", "alog": "_log", "sqrt": "_sqrt", "tanh": "_tanh",
          "acos": "_acos", "asin": "_asin", "tan": "_tan"}


def tokenize(s):
    out, pos = [], 0
    s = s.strip()
    while pos < len(s):
        m = _TOK.match(s, pos)
        if not m:
            raise NotImplementedError(f"cannot tokenize: {s[pos:pos + 40]!r}")
        out.append(m.group(1))
        pos = m.end()
    return out


class Expr:
    """recursive-descent translation of one Fortran expression to Python source, with its static type ('i', 'r', 'd', 'l')"""

    def __init__(self, toks, tr):
        self.t, self.p, self.tr = toks, 0, tr

    def peek(self):
        return self.t[self.p] if self.p < len(self.t) else None

    def take(self, want=None):
        tok = self.peek()
        if want is not None and tok != want:
            raise NotImplementedError(f"expected {want!r}, found {tok!r} in {' '.join(self.t)}")
        self.p += 1
        return tok

    def parse(self):
        r = self.eqv()
        if self.peek() is not None:
            raise NotImplementedError(f"trailing tokens {self.t[self.p:]} in {' '.join(self.t)}")
        return r

    def eqv(self):
        a, ta = self.or_()
        while self.peek() in (".eqv.", ".neqv."):
            op = "==" if self.take() == ".eqv." else "!="
            b, _ = self.or_()
            a, ta = f"(bool({a}) {op} bool({b}))", "l"
        return a, ta

    def or_(self):
        a, ta = self.and_()
        while self.peek() == ".or.":
            self.take()
            b, _ = self.and_()
            a, ta = f"({a} or {b})", "l"
        return a, ta

    def and_(self):
        a, ta = self.not_()
        while self.peek() == ".and.":
            self.take()
            b, _ = self.not_()
            a, ta = f"({a} and {b})", "l"
        return a, ta

    def not_(self):
        if self.peek() == ".not.":
            self.take()
            a, _ = self.not_()
            return f"(not {a})", "l"
        return self.rel()

    def rel(self):
        a, ta = self.add()
        if self.peek() in _REL:
            op = _REL[self.take()]
            b, tb = self.add()
            a, b, _ = self.tr.promote(a, ta, b, tb)
            return f"({a} {op} {b})", "l"
        return a, ta

    def add(self):
        if self.peek() in ("-", "+"):
            op = self.take()
            a, ta = self.mul()
            if op == "-":
                a = f"(-{a})"
        else:
            a, ta = self.mul()
        while self.peek() in ("+", "-"):
            op = self.take()
            b, tb = self.mul()
            a, ta = self.tr.binop(op, a, ta, b, tb)
        return a, ta

    def mul(self):
        a, ta = self.pow_()
        while self.peek() in ("*", "/"):
            op = self.take()
            b, tb = self.pow_()
            a, ta = self.tr.binop(op, a, ta, b, tb)
        return a, ta

    def pow_(self):
        a, ta = self.unary()
        if self.peek() == "**":
            self.take()
            b, tb = self.pow_unary()
            if tb == "i":
                return (f"_powi({a}, {b})", "r") if ta in "rd" else (f"({a} ** {b})", "i")
            return f"_powf({a}, {b})", "r"
        return a, ta

    def pow_unary(self):          # the exponent may carry a sign: x**-2 is not standard, x**(-2) is; handle both
        if self.peek() == "-":
            self.take()
            a, ta = self.pow_()
            return f"(-{a})", ta
        return self.pow_()

    def unary(self):
        if self.peek() == "-":
            self.take()
            a, ta = self.unary()
            return f"(-{a})", ta
        if self.peek() == "+":
            self.take()
            return self.unary()
        return self.atom()

    def args(self):
        self.take("(")
        out = []
        if self.peek() != ")":
            out.append(self.or_())
            while self.peek() == ",":
                self.take()
                out.append(self.or_())
        self.take(")")
        return out

    def atom(self):
        tok = self.take()
        if tok == "(":
            a, ta = self.eqv()
            self.take(")")
            return f"({a})", ta
        if tok[0] in "'\"":
            return repr(tok[1:-1].strip()), "c"
        if tok in (".true.", ".false."):
            return ("True" if tok == ".true." else "False"), "l"
        if re.match(r"[\d.]", tok) and tok[0] != "." or re.match(r"\.\d", tok):
            if re.fullmatch(r"\d+", tok):
                return tok, "i"
            if "d" in tok:
                return f"f64({tok.replace('d', 'e')})", "d"
            t = tok if not tok.endswith(".") else tok + "0"
            return f"f32({t})", "r"
        if not re.match(r"[a-z_]", tok):
            raise NotImplementedError(f"unexpected token {tok!r} in {' '.join(self.t)}")
        return self.tr.resolve(tok, self)


# ---------------------------------------------------------------------------------------------------------------------
class Routine:
    """one translated SUBROUTINE: call it with keyword arguments named like the dummy arguments (arrays: numpy float32 /
    int arrays indexed in Fortran order, first index fastest in the source = first axis here); INTENT(OUT / INOUT) arrays
    are modified in place, scalars that are assigned come back in the returned dict"""

    def __init__(self, path, name, defines=None, types=None, constants=None, library=None):
        self.path, self.name = path, name
        self.library = library             # Library: resolves CALLs to other routines of the reference (or to Python hooks)
        self.types = types or {}           # derived-type argument name -> dict of components given at call time
        self.constants = constants or {}   # USEd module PARAMETERs: name -> (value, type letter), see module_parameters()
        st = statements(path, name, defines, library.texts.get(name) if library is not None else None)
        m = re.match(r"((real|integer|logical)\s+)?(subroutine|function)\s+\w+\s*\((.*)\)\s*$", st[0])
        self.dummies = [a.strip() for a in m.group(4).split(",")] if m and m.group(4).strip() else []
        self.is_function = bool(m) and m.group(3) == "function"
        self.result_kind = {"real": "r", "integer": "i", "logical": "l", None: None}[m.group(2) if m else None]
        self.kind, self.dims, self.params, self.stmtfn, self.derived, self.optional = {}, {}, {}, {}, set(), set()
        self.implicit_none = False
        body = st[1:-1]
        if self.is_function and self.result_kind:
            self.kind[name] = self.result_kind
        if library is not None and name in library.inject:      # what a generated #include declares (dummy_new_args.inc: moist, ...)
            extra_dummies, extra_decl = library.inject[name]
            self.dummies += list(extra_dummies)
            for d in extra_decl:
                assert self._declaration(d.lower()), d
        n = 0
        while n < len(body) and self._declaration(body[n]):
            n += 1
        rest = []                      # declarations may follow statement-function definitions (advect_u: LOGICAL :: specified)
        for st in body[n:]:
            if "::" in st and re.match(r"(real|integer|logical|double precision)\b", st) and self._declaration(st):
                continue
            rest.append(st)
        try:
            self.py = self._translate(rest)
        except NotImplementedError as e:
            if "[in " not in str(e):
                raise NotImplementedError(f"{e} [in {name} of {os.path.basename(path)}]") from e
            raise
        self.fn = None

    # -- declarations
    def _declaration(self, s):
        if s.startswith("implicit"):
            self.implicit_none = True
            return True
        if s.startswith(("use ", "external", "intrinsic", "save")):
            return True
        if re.match(r"character\s*\*\s*\d+", s):        # message buffers: never part of the arithmetic
            return True
        mp = re.match(r"parameter\s*\((.*)\)\s*$", s)     # old style: INTEGER n / PARAMETER (n = 1000)
        if mp:
            for ent in self._split(mp.group(1)):
                nm, init = [x.strip() for x in ent.split("=", 1)]
                if nm not in self.kind:
                    self.kind[nm] = "i" if nm[0] in "ijklmn" else "r"
                self.params[nm] = init
            return True
        m = None
        mt = re.match(r"(real|integer|logical|double precision|character|type\s*\(\s*\w+\s*\))\s*(\([^)]*\))?\s*(,|::)", s)
        if mt and "::" in s:
            sep = s.index("::")
            m = re.match(r"(.*)$", s)      # placeholder match object; the pieces are taken below
            base, attrs, ents = mt.group(1), s[mt.end(3) - len(mt.group(3)):sep], s[sep + 2:]
        if not m:
            m2 = re.match(r"(real|integer|logical|type\s*\(\s*\w+\s*\))\s+([a-z_].*)$", s)        # old style: REAL a, b
            if not m2:
                return False
            base, attrs, ents = m2.group(1), "", m2.group(2)
        k = {"real": "r", "integer": "i", "logical": "l", "double precision": "d", "character": "c"}.get(base.split("(")[0].strip(), "t")
        if k == "r" and mt and mt.group(2) and re.fullmatch(r"\(\s*(kind\s*=\s*)?8\s*\)", mt.group(2)):
            k = "d"                     # REAL(KIND=8)
        dim = None
        md = re.search(r"dimension\s*\(", attrs)
        if md:
            depth, j = 0, md.end() - 1
            for j in range(md.end() - 1, len(attrs)):
                depth += attrs[j] == "("
                depth -= attrs[j] == ")"
                if depth == 0:
                    break
            dim = attrs[md.end():j]
        is_param = "parameter" in attrs
        for ent in self._split(ents):
            ent = ent.strip()
            if not ent:
                continue
            init = None
            if "=" in ent and (is_param or "(" not in ent.split("=", 1)[0]):       # PARAMETER, or an initialised (SAVEd) local
                ent, init = [x.strip() for x in ent.split("=", 1)]
            me = re.match(r"([a-z_]\w*)\s*(\((.*)\))?$", ent)
            if not me:
                raise NotImplementedError(f"declaration entity {ent!r} in {s!r}")
            nm = me.group(1)
            self.kind[nm] = k
            if k == "t":
                self.derived.add(nm)
            if "optional" in attrs:
                self.optional.add(nm)
            d = me.group(3) if me.group(2) else dim
            if d is not None:
                self.dims[nm] = self._split(d)
            if init is not None:
                self.params[nm] = init
        return True

    @staticmethod
    def _split(s):
        out, depth, cur = [], 0, ""
        for ch in s:
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            if ch == "," and depth == 0:
                out.append(cur)
                cur = ""
            else:
                cur += ch
        out.append(cur)
        return [x.strip() for x in out]

    # -- typing helpers used by Expr
    def promote(self, a, ta, b, tb):
        if ta == tb:
            return a, b, ta
        if "d" in (ta, tb):
            return (f"f64({a})" if ta != "d" else a), (f"f64({b})" if tb != "d" else b), "d"
        if "r" in (ta, tb):
            return (f"f32({a})" if ta != "r" else a), (f"f32({b})" if tb != "r" else b), "r"
        return a, b, ta

    def binop(self, op, a, ta, b, tb):
        a, b, t = self.promote(a, ta, b, tb)
        if t == "i":
            return (f"_idiv({a}, {b})" if op == "/" else f"({a} {op} {b})"), "i"
        return f"{'f32' if t == 'r' else 'f64'}({a} {op} {b})", t

    def expr(self, text):
        return Expr(tokenize(text), self).parse()

    def conv(self, src, ts, tt):
        if ts == tt or tt in "lc":
            return src
        if tt == "i":
            return f"int({src})"
        return f"{'f32' if tt == 'r' else 'f64'}({src})"

    def index(self, nm, subs):
        lows = []
        for d in self.dims[nm]:
            lo = d.split(":")[0] if ":" in d else "1"
            lows.append(self.expr(lo)[0])
        if len(subs) != len(lows):
            raise NotImplementedError(f"rank of {nm}")
        return f"v_{nm}[" + ", ".join(f"{s} - {lo}" for s, lo in zip(subs, lows)) + "]"

    def resolve(self, tok, ex):
        if ex.peek() == "%":                          # config_flags%open_xs, grid%u_2(i,k,j), model_config_rec%ideal_case
            ex.take()
            comp = ex.take()
            objects = self.library.objects if self.library is not None else {}
            if tok in objects:
                base = f"v_{tok}" if tok in self.kind else f"_G[{tok!r}]"
                t = objects[tok].get(comp, objects[tok].get("*", "r"))
                if ex.peek() == "(":
                    subs = [a for a, _ in ex.args()]
                    return f"_oget({base}, {comp!r}, ({', '.join(subs)},))", t
                return f"_oget({base}, {comp!r}, None)", t
            return f"v_{tok}[{comp!r}]", self.types.get(tok, {}).get(comp, "l")
        if ex.peek() == "(":
            if tok in self.dims:
                subs = [a for a, _ in ex.args()]
                return self.index(tok, subs), self.kind[tok]
            if tok in self.stmtfn:
                args = ex.args()
                fargs, _ = self.stmtfn[tok]
                conv = [self.conv(a, t, self.kind.get(fa, "r")) for (a, t), fa in zip(args, fargs)]
                return f"sf_{tok}({', '.join(conv)})", self.kind.get(tok, "r")
            args = ex.args()
            return self.intrinsic(tok, args)
        if tok in self.params and tok not in self.kind:
            raise NotImplementedError(tok)
        if tok not in self.kind and tok in self.constants:
            v, t = self.constants[tok]
            return (f"f32({float(v)!r})" if t == "r" else repr(int(v))), t
        if tok not in self.kind and self.library is not None and tok in self.library.globals:
            return f"_G[{tok!r}]", "i"            # a USEd module variable (p_qv, ...), looked up when the routine runs
        if tok not in self.kind:
            raise NotImplementedError(f"undeclared name {tok!r} (IMPLICIT NONE expected)")
        return f"v_{tok}", self.kind[tok]

    def intrinsic(self, fn, args):
        src = [a for a, _ in args]
        ty = [t for _, t in args]
        if fn in ("max", "min", "amax1", "amin1", "max0", "min0"):
            base = "max" if "max" in fn else "min"
            t = "d" if "d" in ty else ("r" if "r" in ty else "i")
            c = [self.conv(a, ta, t) for a, ta in args]
            return f"{base}({', '.join(c)})", t
        if fn == "abs":
            return f"abs({src[0]})", ty[0]
        if fn == "sign":
            t = ty[0] if ty[0] != "i" or ty[1] == "i" else ty[0]
            out = f"_sign({src[0]}, {src[1]})"
            return (f"f32({out})" if t == "r" else out), t
        if fn in ("real", "float", "sngl"):
            return f"f32({src[0]})", "r"
        if fn in ("int", "ifix"):
            return f"int({src[0]})", "i"
        if fn == "nint":
            return f"int(np.round({src[0]}))", "i"
        if fn == "mod":
            return (f"int(np.fmod({src[0]}, {src[1]}))", "i") if ty[0] == "i" else (f"f32(np.fmod({src[0]}, {src[1]}))", "r")
        if fn == "dble":
            return f"f64({src[0]})", "d"
        if fn == "dim":                 # positive difference
            t = "r" if "r" in ty else "i"
            a, b = self.conv(src[0], ty[0], t), self.conv(src[1], ty[1], t)
            return (f"_dim({a}, {b})", t)
        if fn == "ichar":
            return f"ord({src[0]})", "i"
        if fn == "char":
            return f"chr({src[0]})", "c"
        if fn == "trim":
            return f"({src[0]}).rstrip()", "c"
        if fn == "present":
            return f"({src[0]} is not None)", "l"
        if fn in _INTR1:
            return f"{_INTR1[fn]}({self.conv(src[0], ty[0], 'r')})", "r"
        if self.library is not None and fn in self.library.hooks:       # an EXTERNAL function supplied by the test (wrf_dm_on_monitor)
            return f"_hookf({fn!r}, [{', '.join(src)}])", self.kind.get(fn, "r")
        if self.library is not None and fn in self.library.where:
            callee = self.library.routine(fn)
            if callee.is_function:
                return f"_callf({fn!r}, [{', '.join(src)}])", callee.kind.get(fn, "r")
        raise NotImplementedError(f"function or undeclared array {fn!r}")

    # -- statements
    def _translate(self, body):
        out = ["def _routine(A):"]
        ind = 1

        def emit(s):
            out.append("    " * ind + s)

        for nm in self.dummies:
            emit(f"v_{nm} = A[{nm!r}]")
        for nm in self.dummies:          # sequence association: a 2-D actual for a (ims:ime, 1:1, jms:jme) dummy (calculate_full of mu)
            if nm in self.dims and all(d.strip() not in (":", "*") for d in self.dims[nm]):
                shape = []
                for d in self.dims[nm]:
                    lo, hi = (d.split(":") + ["1"])[:2] if ":" in d else ("1", d)
                    shape.append(f"({self.expr(hi)[0]}) - ({self.expr(lo)[0]}) + 1")
                emit(f"v_{nm} = _fit(v_{nm}, {len(shape)}, lambda: ({', '.join(shape)},))")
        for nm, init in self.params.items():
            src, t = self.expr(init)
            emit(f"v_{nm} = {self.conv(src, t, self.kind[nm])}")
        for nm, k in self.kind.items():
            if nm in self.dummies or nm in self.params:
                continue
            if k == "t":
                emit(f"v_{nm} = _Obj()")
                continue
            if k == "c":
                emit(f"v_{nm} = ''")
                continue
            if nm in self.dims:
                shape = []
                for d in self.dims[nm]:
                    lo, hi = (d.split(":") + ["1"])[:2] if ":" in d else ("1", d)
                    shape.append(f"({self.expr(hi)[0]}) - ({self.expr(lo)[0]}) + 1")
                dt = {"r": "np.float32", "i": "np.int64", "l": "np.bool_", "d": "np.float64"}[k]
                fill = "np.nan" if k in "rd" else "0"
                emit(f"v_{nm} = np.full(({', '.join(shape)},), {fill}, dtype={dt})")
            else:
                emit(f"v_{nm} = None")
        blocks = []          # ('do', label) / ('if',)
        n = 0
        while n < len(body):
            s = body[n]
            n += 1
            m = re.match(r"([a-z_]\w*)\s*:\s*(do\b.*|if\b.*|select\b.*)$", s)
            label = None
            if m and not re.match(r"[a-z_]\w*\s*:\s*:", s):
                label, s = m.group(1), m.group(2)
            if re.match(r"(end\s*do|enddo)\b", s):
                ind -= 1
                blocks.pop()
                continue
            msel = re.match(r"select\s*case\s*\(", s)
            if msel:                    # SELECT CASE -> an if / elif chain on the selector's value
                nsel = sum(1 for b in blocks if b[0] == "select")
                emit(f"_sel{nsel} = {self.expr(self._paren_after(s, 6))[0]}")
                emit("if False:")
                ind += 1
                emit("pass")
                blocks.append(("select", label))
                continue
            if re.match(r"case\s+default\b", s):
                ind -= 1
                emit("else:")
                ind += 1
                emit("pass")
                continue
            if re.match(r"case\s*\(", s):
                nsel = sum(1 for b in blocks if b[0] == "select") - 1
                tests = []
                for v in self._split(self._paren_after(s, 4)):
                    if ":" in v:
                        lo, hi = [t.strip() for t in v.split(":")]
                        tests.append(" and ".join(([f"_sel{nsel} >= {self.expr(lo)[0]}"] if lo else []) + ([f"_sel{nsel} <= {self.expr(hi)[0]}"] if hi else [])))
                    else:
                        tests.append(f"_sel{nsel} == {self.expr(v)[0]}")
                ind -= 1
                emit("elif " + " or ".join(f"({t})" for t in tests) + ":")
                ind += 1
                emit("pass")
                continue
            if re.match(r"end\s*select\b", s):
                ind -= 1
                blocks.pop()
                continue
            if re.match(r"(end\s*if|endif)\b", s):
                ind -= 1
                blocks.pop()
                continue
            if re.match(r"else\s*if\b", s):
                cond = self._paren_after(s, s.index("if") + 2)
                ind -= 1
                emit(f"elif {self.expr(cond)[0]}:")
                ind += 1
                emit("pass")
                continue
            if re.match(r"else\b", s):
                ind -= 1
                emit("else:")
                ind += 1
                emit("pass")
                continue
            md = re.match(r"do\s+([a-z_]\w*)\s*=\s*(.*)$", s)
            if md:
                parts = self._split(md.group(2))
                lo, hi = self.expr(parts[0])[0], self.expr(parts[1])[0]
                var = md.group(1)
                if len(parts) == 3:
                    step = self.expr(parts[2])[0]
                    emit(f"for v_{var} in range({lo}, ({hi}) + (1 if ({step}) > 0 else -1), {step}):")
                else:
                    emit(f"for v_{var} in range({lo}, ({hi}) + 1):")
                ind += 1
                emit("pass")
                blocks.append(("do", label))
                continue
            mw = re.match(r"do\s*while\s*\(", s)
            if mw:
                emit(f"while {self.expr(self._paren_after(s, 2))[0]}:")
                ind += 1
                emit("pass")
                blocks.append(("do", label))
                continue
            if re.match(r"do\s*$", s):
                raise NotImplementedError(s)
            if s.startswith("if"):
                cond = self._paren_after(s, 2)
                rest = s[s.index(cond) + len(cond) + 1:].strip()
                csrc = self.expr(cond)[0]
                if rest == "then":
                    emit(f"if {csrc}:")
                    ind += 1
                    emit("pass")
                    blocks.append(("if", label))
                else:
                    emit(f"if {csrc}:")
                    ind += 1
                    self._simple(rest, emit)
                    ind -= 1
                continue
            self._simple(s, emit)
        emit(self._return())
        return "\n".join(out)

    def _return(self):
        if self.is_function:
            return f"return v_{self.name}"
        ret = [nm for nm in self.dummies if nm not in self.dims and self.kind.get(nm) in ("r", "i", "l", "d")]
        return "return {" + ", ".join(f"{nm!r}: v_{nm}" for nm in ret) + "}"

    def _array_dummy(self, callee, key, position):
        """does the callee declare the dummy this actual lands on as an array?"""
        if self.library is None or callee not in self.library.where:
            return False
        R = self.library.routine(callee)
        dummy = key if key else (R.dummies[position] if position < len(R.dummies) else None)
        return dummy in R.dims

    @staticmethod
    def _paren_after(s, pos):
        start = s.index("(", pos)
        depth = 0
        for j in range(start, len(s)):
            depth += s[j] == "("
            depth -= s[j] == ")"
            if depth == 0:
                return s[start + 1:j]
        raise NotImplementedError(s)

    def _simple(self, s, emit):
        if s in ("return",):
            emit(self._return())
            return
        if s == "cycle":
            emit("continue")
            return
        if s == "exit":
            emit("break")
            return
        if s == "continue" or re.match(r"(write\s*\(|print\s*\*)", s):      # diagnostics text is not part of the arithmetic
            emit("pass")
            return
        if s.startswith("call "):
            if re.match(r"call\s+(wrf_debug|wrf_message)\b", s):
                emit("pass")
                return
            mv = re.match(r"call\s+(vpow|vspow)\s*\((.*)\)\s*$", s)
            if mv:          # frame/libmassv.F:375-381 (VPOW is vspow unless DOUBLE_PRECISION): z(j) = y(j)**x(j), j = 1..n, on array
                z, y, x, cnt = self._split(mv.group(2))          # elements passed as the start of a run along the first dimension
                def run_of(text):
                    me = re.match(r"([a-z_]\w*)\s*\((.*)\)$", text.strip())
                    subs = [self.expr(t)[0] for t in self._split(me.group(2))]
                    subs[0] = f"({subs[0]}) + _n"
                    return self.index(me.group(1), subs)
                emit(f"for _n in range({self.expr(cnt)[0]}):")
                emit(f"    {run_of(z)} = _powf({run_of(y)}, {run_of(x)})")
                return
            if re.match(r"call\s+wrf_error_fatal3?\b", s):
                emit(f"raise RuntimeError({s!r})")
                return
            mc = re.match(r"call\s+([a-z_]\w*)\s*(\((.*)\))?\s*$", s)
            if not mc or self.library is None:
                raise NotImplementedError(s)
            pos, kw, back = [], [], []
            for n_arg, a in enumerate(self._split(mc.group(3)) if mc.group(3) and mc.group(3).strip() else []):
                mk = re.match(r"([a-z_]\w*)\s*=\s*([^=].*)$", a)
                key, text = (mk.group(1), mk.group(2).strip()) if mk else (None, a.strip())
                me = re.match(r"([a-z_]\w*)\s*\((.*)\)$", text)
                if me and me.group(1) in self.dims and mc.group(1) not in self.library.hooks and self._array_dummy(mc.group(1), key, n_arg):
                    # an element as the start of a slab (sequence association): moist(ims,kms,jms,im) -> the species' 3-D array
                    nm_a, subs = me.group(1), [t.strip() for t in self._split(me.group(2))]
                    lows = [(d.split(":")[0].strip() if ":" in d else "1") for d in self.dims[nm_a]]
                    lead = 0
                    while lead < len(subs) and subs[lead] == lows[lead]:
                        lead += 1
                    if lead == 0 or len(subs) != len(lows):
                        raise NotImplementedError(f"array element actual {text!r}")
                    idx = [":"] * lead + [f"({self.expr(t)[0]}) - ({self.expr(lo)[0]})" for t, lo in zip(subs[lead:], lows[lead:])]
                    src = f"v_{nm_a}[{', '.join(idx)}]"
                elif re.fullmatch(r"[a-z_]\w*", text) and text in self.kind:      # a variable: arrays and derived types go by reference
                    src = f"v_{text}"
                    if text not in self.dims and self.kind[text] in "ridl":
                        back.append((key if key else n_arg, text))
                else:
                    src = self.expr(text)[0]
                (kw if key else pos).append((key, src))
            emit(f"_r = _call({mc.group(1)!r}, [{', '.join(x for _, x in pos)}], {{{', '.join(f'{k!r}: {x}' for k, x in kw)}}})")
            for key, var in back:
                emit(f"v_{var} = _r.get({key!r}, v_{var})")
            return
        # assignment (or a statement function definition)
        depth, eq = 0, -1
        for j, ch in enumerate(s):
            depth += ch == "("
            depth -= ch == ")"
            if ch == "=" and depth == 0 and s[j - 1] not in "<>/=" and s[j + 1:j + 2] != "=":
                eq = j
                break
        if eq < 0:
            raise NotImplementedError(f"statement {s!r}")
        lhs, rhs = s[:eq].strip(), s[eq + 1:].strip()
        mo = re.match(r"([a-z_]\w*)\s*%\s*([a-z_]\w*)\s*(\((.*)\))?$", lhs)
        if mo and self.library is not None and mo.group(1) in self.library.objects:
            obj, comp = mo.group(1), mo.group(2)
            table = self.library.objects[obj]
            src, t = self.expr(rhs)
            val = self.conv(src, t, table.get(comp, table.get("*", "r")))
            base = f"v_{obj}" if obj in self.kind else f"_G[{obj!r}]"
            subs = "None" if not mo.group(3) else "(" + ", ".join(self.expr(x)[0] for x in self._split(mo.group(4))) + ",)"
            emit(f"_oset({base}, {comp!r}, {subs}, {val})")
            return
        ml = re.match(r"([a-z_]\w*)\s*(\((.*)\))?$", lhs)
        if not ml:
            raise NotImplementedError(f"left-hand side {lhs!r}")
        nm = ml.group(1)
        if ml.group(2) and nm not in self.dims:       # statement function
            fargs = [a.strip() for a in ml.group(3).split(",")]
            self.stmtfn[nm] = (fargs, rhs)
            saved = {a: self.dims.pop(a) for a in fargs if a in self.dims}
            src, t = self.expr(rhs)
            self.dims.update(saved)
            emit(f"def sf_{nm}({', '.join('v_' + a for a in fargs)}):")
            emit(f"    return {self.conv(src, t, self.kind.get(nm, 'r'))}")
            return
        src, t = self.expr(rhs)
        if nm not in self.kind and not self.implicit_none and not ml.group(2):
            self.kind[nm] = "i" if nm[0] in "ijklmn" else "r"      # Fortran's implicit typing (ww_split's cw_max)
        if nm not in self.kind:
            raise NotImplementedError(f"undeclared {nm!r}")
        val = self.conv(src, t, self.kind[nm])
        if ml.group(2) and all(x.strip() == ":" for x in self._split(ml.group(3))):
            emit(f"v_{nm}[...] = {val}")                 # a(:,:,:) = scalar
        elif ml.group(2):
            subs = [self.expr(x)[0] for x in self._split(ml.group(3))]
            emit(f"{self.index(nm, subs)} = {val}")
        elif nm in self.dims:
            emit(f"v_{nm}[...] = {val}")
        else:
            emit(f"v_{nm} = {val}")

    calls = 0          # how often the routine ran (the inventory of tests/golden/reference_source.json)

    def __call__(self, **args):
        self.calls += 1
        if self.fn is None:
            ns = dict(RUNTIME)
            ns["_G"] = self.library.globals if self.library is not None else {}
            ns["_hookf"] = (lambda nm, a: self.library.hooks[nm](*a)) if self.library is not None else None
            ns["_call"] = self.library.call if self.library is not None else None
            ns["_callf"] = self.library.callf if self.library is not None else None
            exec(compile(self.py, f"<{self.name} of {self.path}>", "exec"), ns)
            self.fn = ns["_routine"]
        missing = [d for d in self.dummies if d not in args and d not in self.optional]
        if missing:
            raise TypeError(f"{self.name}: missing arguments {missing}")
        for d in self.optional:
            args.setdefault(d, None)
        with np.errstate(all="ignore"):
            return self.fn(args)


def module_parameters(path, known=None, header_only=False):
    """the REAL / INTEGER PARAMETERs of a module file (share/module_model_constants.F), evaluated in order with real*4
    semantics: name -> (value, 'r' | 'i').  header_only: stop at CONTAINS (the module-level declarations of a file of
    routines); known: constants the expressions may use"""
    out = dict(known or {})
    mine = {}

    class _M(Routine):
        def __init__(self):          # only the expression machinery
            self.kind, self.dims, self.params, self.stmtfn, self.derived, self.types = {}, {}, {}, {}, set(), {}
            self.constants = out
            self.library = None

    tr = _M()
    for ln in open(path, errors="replace"):
        ln = _strip_comment(ln).strip().lower()
        if header_only and ln == "contains":
            break
        m = re.match(r"(real|integer)\s*,\s*parameter\s*::\s*([a-z_]\w*)\s*=\s*(.+)$", ln)
        if not m:
            continue
        try:
            src, t = tr.expr(m.group(3))
        except (NotImplementedError, IndexError):
            continue
        want = "r" if m.group(1) == "real" else "i"
        val = eval(tr.conv(src, t, want), dict(RUNTIME))
        out[m.group(2)] = mine[m.group(2)] = (val, want)
    return mine


class Library:
    """the routines CALL statements may reach: every SUBROUTINE of the given reference files, translated on first use, plus
    Python hooks for what is not Fortran-subset material (namelist getters, message routines)"""

    def __init__(self, files, defines=None, types=None, constants=None, hooks=None):
        self.where, self.cache = {}, {}
        self.globals = {}              # USEd module variables (integers, or objects such as model_config_rec), by name
        self.texts = {}                # name -> source text of a routine assembled by the test (add_fragment)
        self.inject = {}               # routine -> (extra dummy names, extra declaration statements) standing for a generated #include
        self.objects = {}              # derived-type objects with array components: name -> {component: type letter, "*": default}
        self.defines, self.types, self.constants, self.hooks = defines or {}, types or {}, constants or {}, hooks or {}
        for path in files:
            for ln in open(path, errors="replace"):
                m = re.match(r"\s*(?:(?:real|integer|logical)\s+)?(?:subroutine|function)\s+([a-z_]\w*)", ln, re.I)
                if m:
                    self.where.setdefault(m.group(1).lower(), path)

    def add_fragment(self, name, path, first, last, dummies, declarations, trim=0):
        """a line range of a reference file (first/last: regular expressions matched against the lines, both inclusive) wrapped as
        SUBROUTINE name(dummies) with the given declarations: a way to execute a block of a routine too large or too entangled to run
        whole (the first-trip diagnosis inside start_domain_em); trim: lines to drop at the end (the ENDIF that closes an enclosing block).
        The text is assembled in memory; nothing is written anywhere."""
        lines = open(path, errors="replace").read().split("\n")
        a = next(n for n, ln in enumerate(lines) if re.search(first, ln))
        b = next(n for n, ln in enumerate(lines) if n > a and re.search(last, ln))
        self.texts[name] = "\n".join([f"subroutine {name}({', '.join(dummies)})", "implicit none"] + list(declarations) + lines[a:b + 1 - trim]
                                     + [f"end subroutine {name}"])
        self.where[name] = path
        self.cache.pop(name, None)
        return a + 1, b + 1

    def routine(self, name):
        if name not in self.cache:
            if name not in self.where:
                raise NotImplementedError(f"CALL {name}: no such routine in the library")
            self.cache[name] = Routine(self.where[name], name, self.defines, self.types, self.constants, self)
        return self.cache[name]

    def callf(self, name, pos):
        R = self.routine(name)
        return R(**dict(zip(R.dummies, pos)))

    def call(self, name, pos, kw):
        """-> {position or keyword of the caller's actual: value of the callee's dummy after the call} (scalars only)"""
        if name in self.hooks:
            return self.hooks[name](*pos, **kw) or {}
        R = self.routine(name)
        args = dict(zip(R.dummies, pos))
        args.update(kw)
        ret = R(**args)
        out = {n: ret[d] for n, d in enumerate(R.dummies[:len(pos)]) if d in ret}
        out.update({k: ret[k] for k in kw if k in ret})
        return out
