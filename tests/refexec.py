"""Executes the reference's OWN statements.

TEST INFRASTRUCTURE ONLY.  No JVM exists in this image, so the Java cannot run as Java.  But the methods that carry the path's
arithmetic are written in a C-like subset of Java (declarations, for / while / do / if, array indexing, field access, calls,
Math.*): the mean-field sweep and the free energy (algorithm/MeanFieldForBayesNet.java), both getLogProbFor drivers
(models/PhyloBayesModel.java, PhyloBackground.java), the substitution tables (evolution/*.java), the Jstacs E-step, strand
mixture, normalisations, numerical gradient, optimiser, termination conditions and start-distance forecaster (libs/jstacs.jar),
the M-step objective and the branch-length objectives (optimizing/*.java), the selector, the alignment-based models and the
classification thresholds.  This module reads those files from /root/reference AT TEST TIME, translates the method bodies token
by token into Python (same statements, same order, same operands; a construct outside the subset raises) and runs them over
small objects that carry the field and method names the Java uses.  scripts/show_reference_translation.py prints what is
generated.  No reference source is copied into the repository: without /root/reference the functions here raise and the tests
that need them are skipped; the golden vectors they produced (tests/golden/refexec_meanfield.npz, scripts/make_refexec_golden.py)
are committed and are what travels to the GPU box.

What is mechanical: every statement of the translated methods.  What is NOT: the object graph the statements walk (tree
structure from the Newick string, parent / child order: oracle/np_restatement.py; pinned separately against the reference's own
sampler output, tests/test_generator_pins.py) and the few shims named in the class docstrings below (the mean-field cache, the
hand-over of new parameters to the tables, java.util collections).
"""
import math
import os
import re

REFERENCE_ROOT = os.environ.get("PHYFOO_REFERENCE_ROOT", "/root/reference")      # (the tests set it to a missing path to rehearse the GPU box)
REFERENCE_FILE = REFERENCE_ROOT + "/main/java/src/algorithm/MeanFieldForBayesNet.java"

_TOKEN = re.compile(r"""
    (?P<bc>/\*.*?\*/) | (?P<lc>//[^\n]*) | (?P<str>"(?:\\.|[^"\\])*"|'(?:\\.|[^'\\])+') |
    (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[dDfFlL]?) |
    (?P<id>[A-Za-z_$][A-Za-z_0-9$]*) |
    (?P<op>\+\+|--|\+=|-=|\*=|/=|%=|==|!=|<=|>=|&&|\|\||[{}()\[\];,.<>=+\-*/%!?:@&|^~]) |
    (?P<ws>\s+)
""", re.S | re.X)

TYPES = {"int", "double", "boolean", "long", "float"}


def _py(name):
    """a Java identifier that is a Python keyword gets a trailing underscore"""
    import keyword
    return name + "_" if keyword.iskeyword(name) else name


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            if ord(src[pos]) > 127:                # a stray non-ASCII byte outside comments: skip
                pos += 1
                continue
            raise SyntaxError(f"cannot tokenize at {src[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind in ("bc", "lc", "ws"):
            continue
        out.append((kind, m.group(kind)))
    return out


def _match(tokens, i, open_, close):
    """index of the token closing the bracket opened at tokens[i]"""
    depth = 0
    for j in range(i, len(tokens)):
        if tokens[j][1] == open_:
            depth += 1
        elif tokens[j][1] == close:
            depth -= 1
            if depth == 0:
                return j
    raise SyntaxError("unbalanced " + open_)


def find_method(tokens, name, n_params, signature=None):
    """(parameter names, body tokens without the outer braces) of the method `name` with n_params parameters (and, when
    overloads share the count, with the parameter list `signature`, e.g. "double[] d, double v, int start, int end")"""
    for i in range(len(tokens) - 2):
        if tokens[i] == ("id", name) and tokens[i + 1][1] == "(" and i > 0 and (tokens[i - 1][0] == "id" or tokens[i - 1][1] == "]"):
            j = _match(tokens, i + 1, "(", ")")
            b = j + 1
            if b < len(tokens) and tokens[b][1] == "throws":
                while tokens[b][1] != "{":
                    b += 1
            if b >= len(tokens) or tokens[b][1] != "{":
                continue                                        # a call, not a declaration
            inner = tokens[i + 2:j]
            params = [t[1] for k, t in enumerate(inner) if t[0] == "id" and (k + 1 == len(inner) or inner[k + 1][1] == ",")]
            if len(params) != n_params:
                continue
            if signature is not None and "".join(t[1] for t in inner) != signature.replace(" ", ""):
                continue
            end = _match(tokens, b, "{", "}")
            return params, tokens[b + 1:end]
    raise LookupError(name)


class Translator:
    """Java-subset statements -> Python source.  Every construct outside the subset raises."""

    def __init__(self, fields, methods, params, int_division=False):
        self.fields, self.methods = set(fields), set(methods)
        self.locals = set(params)
        self.lines = []
        self.int_division = int_division          # a method whose `/` operands are all ints (Java truncates; operands >= 0 here)

    # ---- expressions ---------------------------------------------------------------------------------------------------
    def expr(self, toks):
        depth = 0
        for q, (_, tv) in enumerate(toks):                      # cond ? a : b at the top level of this expression
            depth += tv in ("(", "[") 
            depth -= tv in (")", "]")
            if tv == "?" and depth == 0:
                d2 = 0
                for c in range(q + 1, len(toks)):
                    d2 += toks[c][1] in ("(", "[", "?")
                    d2 -= toks[c][1] in (")", "]")
                    if toks[c][1] == ":" :
                        if d2 == 0:
                            return f"(({self.expr(toks[q + 1:c])}) if ({self.expr(toks[:q])}) else ({self.expr(toks[c + 1:])}))"
                        d2 -= 1
                raise SyntaxError("? without :")
        out, i = [], 0
        while i < len(toks):
            kind, t = toks[i]
            prev = toks[i - 1][1] if i else None
            nxt = toks[i + 1][1] if i + 1 < len(toks) else None
            if t == "(" and i + 1 < len(toks) and toks[i + 1][0] == "id" and toks[i + 1][1][0].isupper():
                j = i + 2                                       # ( Type ) x  or  ( Type < ... > ) x : a class cast -- dropped
                if j < len(toks) and toks[j][1] == "<":
                    j = _match(toks, j, "<", ">") + 1
                if j + 1 < len(toks) and toks[j][1] == ")" and (toks[j + 1][0] == "id" or toks[j + 1][1] == "("):
                    i = j + 1
                    continue
            if t == "instanceof":                               # x instanceof T  (x a single name)
                out[-1] = f'_instanceof({out[-1]}, "{toks[i + 1][1]}")'
                i += 2
                continue
            if t == "new" and toks[i + 1][1] == "ArrayList":    # new ArrayList<T>()
                j = _match(toks, i + 2, "<", ">") if toks[i + 2][1] == "<" else i + 1
                assert toks[j + 1][1] == "(" and toks[j + 2][1] == ")"
                out.append("_JList()")
                i = j + 3
                continue
            if t == "(" and i + 2 < len(toks) and toks[i + 1][1] in TYPES and toks[i + 2][1] == ")":
                if toks[i + 1][1] in ("int", "long") and i + 3 < len(toks) and toks[i + 3][1] == "(":
                    j = _match(toks, i + 3, "(", ")")            # (int) ( expr ): Java truncates towards zero, like int()
                    out.append(f"int({self.expr(toks[i + 4:j])})")
                    i = j + 1
                    continue
                i += 3                                          # a primitive cast in front of a value of that type already
                continue
            if t == "/" and self.int_division:
                out.append("//")
                i += 1
                continue
            if t == "new":                                      # new double[expr] / new int[expr]
                ty = toks[i + 1][1]
                j = _match(toks, i + 2, "[", "]")
                zero = {"double": "0.0", "int": "0", "boolean": "False", "String": "None"}[ty]
                if j + 1 < len(toks) and toks[j + 1][1] == "[":
                    j2 = _match(toks, j + 1, "[", "]")
                    out.append(f"_JArray(_JArray([{zero}] * ({self.expr(toks[j + 2:j2])})) for _ in range({self.expr(toks[i + 3:j])}))")
                    i = j2 + 1
                    continue
                out.append(f"_JArray([{zero}] * ({self.expr(toks[i + 3:j])}))")
                i = j + 1
                continue
            if kind == "id":
                if prev == ".":
                    out.append(t)
                elif t in ("true", "false"):
                    out.append(t.capitalize())
                elif t == "null":
                    out.append("None")
                elif t == "this":
                    out.append("self")
                elif t in self.locals:
                    out.append(_py(t))
                elif t in self.methods and nxt == "(":
                    out.append("self." + t)
                elif t in self.fields:
                    out.append("self." + t)
                elif t in ("Math", "Double", "Alphabet", "MeanFieldForBayesNet", "Arrays", "System", "Normalisation", "SimpleNodeElimination", "PhyloSample", "AlignmentBasedModel", "EdgeOptimizer"):
                    out.append(t)
                else:
                    raise NameError(f"unknown identifier {t!r} in the reference statement")
            elif kind == "num":
                out.append(t.rstrip("dDfFlL"))
            elif t == "&&":
                out.append("and")
            elif t == "||":
                out.append("or")
            elif t == "!":
                out.append("not")
            elif t in ("++", "--"):
                raise NotImplementedError(t + " inside an expression")
            elif kind == "str":
                out.append(t)
            else:
                out.append(t)
            i += 1
        s = " ".join(out)
        for a, b in (("Math . log", "_jlog"), ("Math . exp", "_jexp"), ("Math . abs", "abs"), ("Math . min", "min"),
                     ("Math . max", "max"), ("Arrays . copyOf", "_copyOf"), ("Double . isNaN", "_isnan"), ("Arrays . sort", "_sort"), ("Math . sqrt", "_sqrt"), ("Double . POSITIVE_INFINITY", "_POS_INF"), ("Double . isInfinite", "_isinf"), ("Normalisation .", "_Normalisation ."),
                     ("Double . NEGATIVE_INFINITY", "_NEG_INF"), ("Alphabet . size", "4"),
                     ("MeanFieldForBayesNet . CALC_LIKELIHOOD", "self.CALC_LIKELIHOOD")):
            s = s.replace(a, b)
        return s

    # ---- statements ----------------------------------------------------------------------------------------------------
    def emit(self, ind, text):
        self.lines.append("    " * ind + text)

    def split_top(self, toks, sep):
        parts, depth, cur = [], 0, []
        for t in toks:
            if t[1] in "([{":
                depth += 1
            elif t[1] in ")]}":
                depth -= 1
            if t[1] == sep and depth == 0:
                parts.append(cur)
                cur = []
            else:
                cur.append(t)
        parts.append(cur)
        return parts

    def simple(self, toks, ind):
        """one statement without its terminating semicolon"""
        if not toks:
            return
        v = [t[1] for t in toks]
        if v[0] == "continue":
            self.emit(ind, "continue")                          # (labels are checked in stmt())
            return
        if v[0] == "break":
            self.emit(ind, "break")
            return
        if v[0] == "return":
            self.emit(ind, "return " + self.expr(toks[1:]) if len(toks) > 1 else "return")
            return
        if v[:3] == ["System", ".", "out"]:
            self.emit(ind, "pass")
            return
        if v[0] == "myOut":                                     # progress output: only the side effects ( ++i ) are kept
            bumped = [v[q + 1] for q in range(len(v) - 1) if v[q] == "++" and toks[q + 1][0] == "id"]
            for name in bumped:
                self.emit(ind, f"{name} += 1")
            if not bumped:
                self.emit(ind, "pass")
            return
        if v[0] == "throw":
            self.emit(ind, "raise RuntimeError('the reference throws here')")
            return
        eq = next((k for k, x in enumerate(v) if x in ("=", "+=", "-=", "*=", "/=", "%=")), None)
        if eq is not None and "?" in v[eq + 1:] and not any(x in ("++", "--") for x in v):
            self.emit(ind, f"{self.expr(toks[:eq])} {v[eq]} {self.expr(toks[eq + 1:])}")
            return
        # NAME [ VAR ++ ] = rhs   with rhs free of VAR:   NAME[VAR] = rhs; VAR += 1   (Java takes the index before the increment)
        if eq is not None and eq >= 5 and v[1] == "[" and v[3] == "++" and v[4] == "]" and eq == 5 and v[eq] == "=" \
                and toks[2][0] == "id" and v[2] not in v[eq + 1:] and not any(x in ("++", "--") for x in v[eq + 1:]):
            self.emit(ind, f"{self.expr(toks[:1])} [ {self.expr(toks[2:3])} ] = {self.expr(toks[eq + 1:])}")
            self.emit(ind, f"{self.expr(toks[2:3])} += 1")
            return
        # post-increments used as values inside the statement (a[i++], w[k++] on either side), each variable occurring once:
        # the statement with the plain variable, then the increments (Java reads the value before incrementing)
        posts = [q for q in range(1, len(v) - 1) if v[q] == "++" and toks[q - 1][0] == "id" and v[q - 1] not in ("]", ")")]
        if posts and all(v.count(v[q - 1]) == 1 for q in posts) and not any(x == "--" for x in v) \
                and not any(v[q] == "++" and q not in posts for q in range(len(v) - 1)):
            rest = [tk for q, tk in enumerate(toks) if q not in posts]
            self.simple(rest, ind)
            for q in posts:
                self.emit(ind, f"{self.expr(toks[q - 1:q])} += 1")
            return
        if any(x in ("++", "--") for x in v[:-1]):              # e.g. a[i++] = ...: outside the subset; fails only if reached
            self.emit(ind, "raise NotImplementedError('statement outside the translated subset of Java')")
            return
        if v[:3] == ["Arrays", ".", "fill"]:
            a, b = self.split_top(toks[4:-1], ",")
            self.emit(ind, f"_fill({self.expr(a)}, {self.expr(b)})")
            return
        # declaration:  type [ [] ]* name [= expr] {, name [= expr]}
        if toks[0][0] == "id" and (v[0] in TYPES or (v[0][0].isupper() and len(toks) > 1 and (toks[1][0] == "id" or v[1] in ("[", "<")))) \
                and v[0] not in ("Math", "System", "Arrays"):
            j = 1
            if j < len(toks) and v[j] == "<" and v[0][0].isupper():          # Type<...> name
                j = _match(toks, j, "<", ">") + 1
            while j + 1 < len(toks) and v[j] == "[" and v[j + 1] == "]":
                j += 2
            if j < len(toks) and toks[j][0] == "id":
                for d in self.split_top(toks[j:], ","):
                    name = d[0][1]
                    self.locals.add(name)
                    if len(d) > 1 and v[0] == "SafeOutputStream":
                        self.emit(ind, f"{_py(name)} = None")
                    elif len(d) > 1 and d[2][1] == "{":          # array initialiser { a, b, ... }
                        assert d[1][1] == "=" and d[-1][1] == "}", d
                        items = ", ".join(self.expr(part) for part in self.split_top(d[3:-1], ","))
                        self.emit(ind, f"{_py(name)} = _JArray([{items}])")
                    elif len(d) > 1:
                        assert d[1][1] == "=", d
                        self.emit(ind, f"{_py(name)} = {self.expr(d[2:])}")
                    else:
                        self.emit(ind, f"{_py(name)} = None")
                return
        if v[-1] in ("++", "--"):
            self.emit(ind, f"{self.expr(toks[:-1])} {'+' if v[-1] == '++' else '-'}= 1")
            return
        self.emit(ind, self.expr(toks))                         # assignment or call

    def stmt(self, toks, i, ind, loop_labels):
        """translates the statement starting at toks[i]; returns the index behind it"""
        t = toks[i][1]
        if t == "{":
            j = _match(toks, i, "{", "}")
            n = len(self.lines)
            k = i + 1
            while k < j:
                k = self.stmt(toks, k, ind, loop_labels)
            if len(self.lines) == n:
                self.emit(ind, "pass")
            return j + 1
        if toks[i][0] == "id" and i + 1 < len(toks) and toks[i + 1][1] == ":" and t not in ("default", "case"):
            return self._labelled(toks, i, ind, loop_labels)
        if t == "do":
            end = _match(toks, i + 1, "{", "}")
            assert toks[end + 1][1] == "while"
            c_end = _match(toks, end + 2, "(", ")")
            self.emit(ind, "while True:")
            self.stmt(toks, i + 1, ind + 1, loop_labels + [None])
            self.emit(ind + 1, f"if not ({self.expr(toks[end + 3:c_end])}):")
            self.emit(ind + 2, "break")
            return c_end + 2                                    # behind the semicolon
        if t in ("for", "while"):
            return self._loop(toks, i, ind, loop_labels, None)
        if t == "try":                                          # try { A } catch (E e) { B }: A runs; B only on an exception
            k = self.stmt(toks, i + 1, ind, loop_labels)
            while k < len(toks) and toks[k][1] in ("catch", "finally"):
                if toks[k][1] == "catch":
                    k = _match(toks, k + 1, "(", ")") + 1
                else:
                    raise NotImplementedError("finally")
                k = _match(toks, k, "{", "}") + 1
            return k
        if t == "if":
            j = _match(toks, i + 1, "(", ")")
            self.emit(ind, f"if {self.expr(toks[i + 2:j])}:")
            k = self.stmt(toks, j + 1, ind + 1, loop_labels)
            if k < len(toks) and toks[k][1] == "else":
                self.emit(ind, "else:")
                k = self.stmt(toks, k + 1, ind + 1, loop_labels)
            return k
        if t == ";":
            return i + 1
        j = i
        depth = 0
        while toks[j][1] != ";" or depth:
            depth += toks[j][1] in "([{"
            depth -= toks[j][1] in ")]}"
            j += 1
        if t == "continue" and j - i == 2:
            if not loop_labels or loop_labels[-1] != toks[i + 1][1]:
                raise NotImplementedError("continue to an outer label")
            self.simple(toks[i:i + 1], ind)
        else:
            self.simple(toks[i:j], ind)
        return j + 1

    def _labelled(self, toks, i, ind, loop_labels):
        if toks[i + 2][1] not in ("for", "while"):
            raise NotImplementedError("label on a non-loop")
        return self._loop(toks, i + 2, ind, loop_labels, toks[i][1])

    def _loop(self, toks, i, ind, loop_labels, label):
        j = _match(toks, i + 1, "(", ")")
        head = toks[i + 2:j]
        labels = loop_labels + [label]
        if toks[i][1] == "while":
            self.emit(ind, f"while {self.expr(head)}:")
            return self.stmt(toks, j + 1, ind + 1, labels)
        colon = [q for q, tk in enumerate(head) if tk[1] == ":"]
        if colon and not any(tk[1] == ";" for tk in head):      # for (Type name : collection)
            var = head[colon[0] - 1][1]
            self.locals.add(var)
            self.emit(ind, f"for {_py(var)} in {self.expr(head[colon[0] + 1:])}:")
            return self.stmt(toks, j + 1, ind + 1, labels)
        init, cond, upd = self.split_top(head, ";")
        iv = [t[1] for t in init]
        cv = [t[1] for t in cond]
        uv = [t[1] for t in upd]
        # for (int v = A; v < B; v++)  /  v <= B   -- the only shape the three methods use
        if len(iv) >= 4 and iv[0] == "int" and iv[2] == "=" and cv[0] == iv[1] and cv[1] in ("<", "<=") and uv == [iv[1], "++"] \
                and not any(x in ("&&", "||", "?") for x in cv) and iv[1] not in cv[2:] and "," not in iv:
            var = iv[1]
            self.locals.add(var)
            lo, hi = self.expr(init[3:]), self.expr(cond[2:])
            self.emit(ind, f"for {var} in range({lo}, ({hi}){' + 1' if cv[1] == '<=' else ''}):")
            return self.stmt(toks, j + 1, ind + 1, labels)
        # general shape: init statements; while (cond) { body; updates }  -- only for bodies without `continue`
        end = _match(toks, j + 1, "{", "}") if toks[j + 1][1] == "{" else None
        if end is None or any(t[1] == "continue" for t in toks[j + 1:end]):
            raise NotImplementedError("for-loop shape " + " ".join(iv + [";"] + cv + [";"] + uv))
        if init and init[0][1] in TYPES:
            self.simple(init, ind)                              # int i = 0, k = 0
        else:
            for part in self.split_top(init, ","):
                self.simple(part, ind)
        self.emit(ind, f"while {self.expr(cond) if cond else 'True'}:")
        k = self.stmt(toks, j + 1, ind + 1, labels)
        for part in self.split_top(upd, ","):
            self.simple(part, ind + 1)
        return k


def _jlog(x):
    """java.lang.Math.log: -Infinity at 0, NaN below"""
    if x > 0:
        return math.log(x)
    return -math.inf if x == 0 else math.nan


def _jexp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


class _JArray(list):
    """a Java array: a list with a .length"""
    length = property(len)


def _copyOf(a, n):
    """java.util.Arrays.copyOf for numeric arrays"""
    return _JArray((list(a) + [0] * n)[:n])


def _fill(a, v):
    for i in range(len(a)):
        a[i] = v


FIELDS = ["bnh", "q", "seq", "parent", "node2hiddenNodeMap", "hiddenNodes", "ns", "CALC_LIKELIHOOD"]
METHODS = ["init", "optimizeByNormalisation", "calcFreeEnergy", "calcLogLikelihood"]


def translate(path=REFERENCE_FILE):
    """{method name: python source} for init(), optimizeByNormalisation(), calcFreeEnergy(start, end), calcFreeEnergy()."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    with open(path, encoding="latin-1") as f:
        toks = tokenize(f.read())
    out = {}
    for py_name, java_name, n in (("init", "init", 0), ("optimizeByNormalisation", "optimizeByNormalisation", 0),
                                  ("calcFreeEnergy2", "calcFreeEnergy", 2), ("calcFreeEnergy0", "calcFreeEnergy", 0)):
        params, body = find_method(toks, java_name, n)
        tr = Translator(FIELDS, METHODS, params)
        tr.emit(0, f"def {py_name}(self{''.join(', ' + _py(p) for p in params)}):")
        k = 0
        while k < len(body):
            k = tr.stmt(body, k, 1, [])
        out[py_name] = "\n".join(tr.lines) + "\n"
    return out


# ---- the object graph the statements walk (names as in bayesNet/*.java) ---------------------------------------------------
class _Props:
    def __init__(self, leaf):
        self._isObserved, self._observedIndex, self._leaf = False, 0, leaf

    def isPhyloLeaf(self):
        return self._leaf

    def setObservation(self, idx):                              # bayesNet/BayesNetNodeProperties.java:88-99
        if idx >= 4:
            self._isObserved = False
        else:
            self._observedIndex, self._isObserved = idx, True


class _CPF:
    def __init__(self, node, table, independent):
        self._node = node
        self._condProb = [list(map(float, r)) for r in table]
        self._indepedentCondProb = [list(map(float, r)) for r in independent]
        self._size = len(self._condProb)
        npar = node.numberOfParents
        self._lookup = [[(l // 4 ** p) % 4 for p in range(npar)] for l in range(self._size)]       # CPF.java:238-256

    def _independent(self):                                     # CPF.java:146-152,218-224 (ALLOW_RETURN_INDEPENDENT_TRANSITION)
        return self._node.props.isPhyloLeaf() and not self._node.props._isObserved

    def get(self, l, obs):
        return (self._indepedentCondProb if self._independent() else self._condProb)[l][obs]

    def getCondProb(self):
        return self._indepedentCondProb if self._independent() else self._condProb

    def getLookUpPointer(self):
        return self._lookup


class _Node:
    def __init__(self, number, leaf):
        self.nodeNumber, self.props = number, _Props(leaf)
        self.Aparents, self.Achildren = [], []

    numberOfParents = property(lambda self: len(self.Aparents))
    numberOfChilds = property(lambda self: len(self.Achildren))

    def isRoot(self):                                           # bayesNet/BayesNetNode.java:144-146
        return len(self.Aparents) == 0

    def getParentIndex(self, parent):                           # :154-156
        return self.Aparents.index(parent)


class _NodeList:
    def __init__(self, nodes):
        self.nodes, self.numberOfNodes = nodes, len(nodes)

    def getNode(self, i):
        return self.nodes[i]


class _Handler:
    def __init__(self, net, trees):
        self._net, self._myTreesArray, self.motifLength = net, trees, len(trees)

    def getNet(self):
        return self._net


class _Seq:
    def __init__(self, length):
        self._length = length

    def getLength(self):
        return self._length


class ReferenceMeanField:
    """MeanFieldForBayesNet with the reference's own method bodies (translated on construction) over the nodes of an
    oracle.np_restatement.Net (tables built).  score(cols) = what PhyloBayesModel.getLogProbFor does with it
    (models/PhyloBayesModel.java:250-260): initObservation, init, optimizeByNormalisation, -calcFreeEnergy(0, W-1)."""
    CALC_LIKELIHOOD = False
    _code = None

    def __init__(self, net):
        if ReferenceMeanField._code is None:
            ns = {"_jlog": _jlog, "_jexp": _jexp, "_NEG_INF": -math.inf, "_fill": _fill, "_JArray": _JArray}
            src = translate()
            for s in src.values():
                exec(compile(s, REFERENCE_FILE + " (translated)", "exec"), ns)
            ReferenceMeanField._code, ReferenceMeanField.source = ns, src
        self.net = net
        index = {n: i for i, n in enumerate(net.order_nodes)}
        leaves = set(net.leaves)
        nodes = [_Node(i, k in leaves) for i, (t, k) in enumerate(net.order_nodes)]
        for n, i in index.items():
            nodes[i].Aparents = [nodes[index[p]] for p in net.par[n]]
            nodes[i].Achildren = [nodes[index[c]] for c in net.chi[n]]
        for n, i in index.items():
            nodes[i].CPF = _CPF(nodes[i], net.tab[n].reshape(-1, 4), net.ind[n].reshape(-1, 4))
        self.nodes = nodes
        trees = [_NodeList([nodes[index[(t, k)]] for k in range(net.N)]) for t in range(net.W)]
        self.leaf_nodes = [[nodes[index[(t, k)]] for k in net.leaves] for t in range(net.W)]
        self.bnh = _Handler(_NodeList(nodes), trees)
        self.seq = _Seq(net.W)
        self.q = self.node2hiddenNodeMap = self.parent = self.ns = None
        self.hiddenNodes = 0
        self.free_energy_calls = 0

    # the three translated methods
    def init(self):
        return self._code["init"](self)

    def optimizeByNormalisation(self):
        return self._code["optimizeByNormalisation"](self)

    def calcFreeEnergy(self, *a):
        if a:
            return self._code["calcFreeEnergy2"](self, *a)
        self.free_energy_calls += 1                             # the sweep loop calls the no-argument form: once before, once per sweep
        return self._code["calcFreeEnergy0"](self)

    def initObservation(self, cols, revcomp=False):
        """MeanFieldForBayesNet.initObservation :78-89 -> VirtualTree.setObservation :183-198 (leaves in leaf order); the
        reverse strand reads the complemented columns in reverse order (what the reference's reverse-complement Sequence holds)"""
        W = self.net.W
        for t in range(W):
            col = cols[W - 1 - t] if revcomp else cols[t]
            for s, node in enumerate(self.leaf_nodes[t]):
                c = int(col[s])
                node.props.setObservation(3 - c if (revcomp and c < 4) else c)

    def score(self, cols, revcomp=False):
        """(-F, sweeps)"""
        self.initObservation(cols, revcomp)
        self.init()
        self.free_energy_calls = 0
        self.optimizeByNormalisation()
        sweeps = self.free_energy_calls - 1                     # one call before the loop, one per sweep
        return -self.calcFreeEnergy(0, self.net.W - 1), sweeps


# ---- the ZOOPS E-step: Jstacs statements from libs/jstacs.jar ----------------------------------------------------------------
JSTACS_JAR = REFERENCE_ROOT + "/libs/jstacs.jar"
SHM_JAVA = "de/jstacs/models/mixture/motif/SingleHiddenMotifMixture.java"
NORMALISATION_JAVA = "de/jstacs/utils/Normalisation.java"
PRIOR_JAVA = "de/jstacs/models/mixture/motif/positionprior/UniformPositionPrior.java"


def _jar_tokens(member):
    import zipfile
    with zipfile.ZipFile(JSTACS_JAR) as z:
        return tokenize(z.read(member).decode("latin-1"))


def _translate_method(tokens, py_name, java_name, n, fields, methods, signature=None, case=None, int_division=False):
    params, body = find_method(tokens, java_name, n, signature)
    if case is not None:                                        # the body of one `case X:` of the method's switch
        v = [t[1] for t in body]
        a = next(i for i in range(len(v) - 2) if v[i] == "case" and v[i + 1] == case and v[i + 2] == ":") + 3
        b = next(i for i in range(a, len(v)) if v[i] in ("case", "default"))
        body = body[a:b]
    tr = Translator(fields, methods, params, int_division)
    tr.emit(0, f"def {py_name}(self{''.join(', ' + _py(p) for p in params)}):")
    k = 0
    while k < len(body):
        k = tr.stmt(body, k, 1, [])
    return "\n".join(tr.lines) + "\n"


def reference_lookup_table(n_parents):
    """CPF._lookup as CPF.generateLookUpTable builds it (bayesNet/CPF.java:280-289): row l = getQueryByIndex(l, false)
    (:238-256, executed from the reference; its `/` is Java's int division), with _SizeBuffer[i] = 4^i (:52-59)."""
    path = REFERENCE_ROOT + "/main/java/src/bayesNet/CPF.java"
    with open(path, encoding="latin-1") as f:
        toks = tokenize(f.read())
    src = _translate_method(toks, "getQueryByIndex", "getQueryByIndex", 2, ["_lookup", "_myNode", "_SizeBuffer"], [], int_division=True)
    ns = {"_JArray": _JArray}
    exec(compile(src, path, "exec"), ns)

    class _Owner:
        numberOfParents = n_parents

    class _Cpf:
        _myNode, _lookup = _Owner(), None
        _SizeBuffer = _JArray(4 ** i for i in range(n_parents))
    return [list(ns["getQueryByIndex"](_Cpf(), l, False)) for l in range(4 ** n_parents)], src


def translate_estep():
    """{name: python source}: SingleHiddenMotifMixture.getNewWeights and modify (EM branch)
    (jstacs!/de/jstacs/models/mixture/motif/SingleHiddenMotifMixture.java:499-566,585-596), UniformPositionPrior.
    getLogPriorForPositions (:55-61) and the Normalisation methods they reach (jstacs!/de/jstacs/utils/Normalisation.java:
    199-201,241-253,352-376,472-476)."""
    if not os.path.exists(JSTACS_JAR):
        raise FileNotFoundError(JSTACS_JAR)
    shm, nrm, pri = _jar_tokens(SHM_JAVA), _jar_tokens(NORMALISATION_JAVA), _jar_tokens(PRIOR_JAVA)
    shm_fields = ["sample", "refBgSample", "model", "posPrior", "logWeights", "trainOnlyMotifModel", "bgMaxMarkovOrder", "algorithm"]
    shm_methods = ["getMotifLength", "initWithPrior", "modify"]
    nrm_methods = ["logSumNormalisation", "normalisation", "getLogSum"]
    return {
        "getNewWeights": _translate_method(shm, "getNewWeights", "getNewWeights", 3, shm_fields, shm_methods),
        "modify": _translate_method(shm, "modify", "modify", 4, shm_fields, shm_methods, case="EM"),
        "getLogPriorForPositions": _translate_method(pri, "getLogPriorForPositions", "getLogPriorForPositions", 2, ["motifLength"], []),
        "lsn5": _translate_method(nrm, "lsn5", "logSumNormalisation", 5, [], nrm_methods),
        "lsn6": _translate_method(nrm, "lsn6", "logSumNormalisation", 6, [], nrm_methods,
                                  "double[] d, int startD, int endD, double[] secondValues, double[] dest, int startDest"),
        "lsn7": _translate_method(nrm, "lsn7", "logSumNormalisation", 7, [], nrm_methods),
        "norm4": _translate_method(nrm, "norm4", "normalisation", 4, [], nrm_methods, "double[] d, double v, int start, int end"),
    }


class _Normalisation:
    """static methods of de.jstacs.utils.Normalisation, overloads picked by the number of arguments like javac does here"""

    def __init__(self, ns):
        self._ns = ns

    def logSumNormalisation(self, *a):
        return self._ns[{5: "lsn5", 6: "lsn6", 7: "lsn7"}[len(a)]](self, *a)

    def normalisation(self, d, v, start, end):
        assert isinstance(start, int) and isinstance(end, int)
        return self._ns["norm4"](self, d, v, start, end)


class _Windows:
    def __init__(self, n):
        self.n = n

    def getNumberOfElements(self):
        return self.n

    def getElementAt(self, i):
        return i


class _Alignments:
    def __init__(self, lengths):
        self.seqs = [_AlignedSeq(i, L) for i, L in enumerate(lengths)]

    def getElementAt(self, i):
        return self.seqs[i]


class _AlignedSeq:
    def __init__(self, index, length):
        self.index, self.length = index, length

    def getLength(self):
        return self.length


class _Scores:
    """a component model as getNewWeights sees it: getLogProbFor(sequence-or-window, start, end)"""

    def __init__(self, fn):
        self.fn = fn

    def getLogProbFor(self, what, start, end):
        return self.fn(what, start, end)


class _Prior:
    def __init__(self, ns, motif_length):
        self._ns, self.motifLength = ns, motif_length

    def getLogPriorForPositions(self, seq_length, *starts):
        return self._ns["getLogPriorForPositions"](self, seq_length, starts)


class ReferenceEStep:
    """SingleHiddenMotifMixture.getNewWeights with the reference's own statements.  The two component models are callables:
    motif(window index) -> log-score of that window, background(alignment index, start, end) -> log-score of the column range
    (end inclusive; 0 for an empty range like models/PhyloBackground.java:242-244)."""
    _ns = None

    def __init__(self, lengths, W, motif, background, zoops, bg_order=0, hyper=(0.0, 0.0)):
        if ReferenceEStep._ns is None:
            ns = {"_jlog": _jlog, "_jexp": _jexp, "_NEG_INF": -math.inf, "_fill": _fill, "_JArray": _JArray, "_isinf": math.isinf}
            src = translate_estep()
            for s in src.values():
                exec(compile(s, JSTACS_JAR + " (translated)", "exec"), ns)
            ns["_Normalisation"] = _Normalisation(ns)
            ReferenceEStep._ns, ReferenceEStep.source = ns, src
        self.W, self.lengths, self.hyper = W, [int(x) for x in lengths], hyper
        self.n_windows = sum(L - W + 1 for L in self.lengths)
        self.sample = [_Windows(self.n_windows), _Alignments(self.lengths)]
        self.refBgSample = list(range(len(self.lengths) + 1))
        self.model = [_Scores(lambda w, s, e: motif(w)), _Scores(lambda seq, s, e: background(seq.index, s, e))]
        self.posPrior = _Prior(self._ns, W)
        with_log = lambda x: math.log(x) if x > 0 else -math.inf
        self.logWeights = _JArray([with_log(zoops), with_log(1.0 - zoops)])
        self.trainOnlyMotifModel, self.bgMaxMarkovOrder, self.algorithm = True, bg_order, "EM"

    def getMotifLength(self, component):
        return self.W

    def initWithPrior(self, w):                                 # jstacs AbstractMixtureModel.java:1098-1100 (System.arraycopy)
        w[0], w[1] = self.hyper

    def modify(self, *a):
        return self._ns["modify"](self, *a)

    def getNewWeights(self, data_weights=None):
        """-> (ll, w[2], seqweights[0])"""
        w = _JArray([0.0, 0.0])
        seqweights = _JArray([_JArray([0.0] * self.n_windows), _JArray([])])
        dw = None if data_weights is None else _JArray([float(x) for x in data_weights])
        ll = self._ns["getNewWeights"](self, dw, w, seqweights)
        return ll, list(w), list(seqweights[0])


# ---- the substitution tables: evolution/FS81alpha.java, FS81beta.java, HKY.java (reinit) -------------------------------------
EVOL_DIR = REFERENCE_ROOT + "/main/java/src/evolution/"


class ReferenceEvolModel:
    """reinit(evolDist) of one of the reference's evolution models with its own statements; the transition matrix has the
    reference's layout [a1 + 4 * context][a2] (evolution/FS81alpha.java:76-92, FS81beta.java:81-103, HKY.java:77-106)."""
    _cache = {}

    def __init__(self, name, pi_rows, hky_field=0.0):
        if name not in self._cache:
            path = EVOL_DIR + name + ".java"
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            with open(path, encoding="latin-1") as f:
                toks = tokenize(f.read())
            fields = ["_dimension", "_alphabetSize", "_transitionMatrix", "_pi", "dimension", "transitionMatrix", "pi", "alpha_beta_ratio"]
            src = _translate_method(toks, "reinit", "reinit", 1, fields, [])
            ns = {"_jlog": _jlog, "_jexp": _jexp, "_NEG_INF": -math.inf, "_JArray": _JArray}
            exec(compile(src, path + " (translated)", "exec"), ns)
            self._cache[name] = (ns["reinit"], src)
        self._reinit, self.source = self._cache[name]
        rows = [[float(x) for x in r] for r in pi_rows]
        self._pi = self.pi = _JArray(_JArray(r) for r in rows)
        self._dimension = self.dimension = len(rows)
        self._alphabetSize = 4
        # HKY.java:32 assigns the constructor argument to itself, so the field keeps its default 0; hky_field is what the field holds
        self.alpha_beta_ratio = hky_field
        self._transitionMatrix = self.transitionMatrix = _JArray(_JArray([0.0] * 4) for _ in range(4 * len(rows)))

    def reinit(self, dist):
        self._reinit(self, _JArray([float(dist)]))
        return [list(r) for r in self._transitionMatrix]


# ---- the M-step objective and its forward-difference gradient (rows a14, a15) --------------------------------------------------
MATRIX_LINEARISATION_JAVA = REFERENCE_ROOT + "/main/java/src/util/MatrixLinearisation.java"
NUMDIFF_JAVA = "de/jstacs/algorithms/optimization/NumericalDifferentiableFunction.java"


class ReferenceObjective:
    """FE_byParameter_ForBayesNodeJstacs.evaluateFunction (optimizing/FE_byParameter_ForBayesNodeJstacs.java:43-53) over windows
    whose mean fields stay fixed (AbstractFreeEnergyOptimizer.OPTIMIZE_MEANFIELDS = false, :36): the statements executed from
    the reference are MatrixLinearisation.lambda2pi (util/MatrixLinearisation.java:15-19) with Normalisation.logSumNormalisation,
    MeanFieldForBayesNet.calcFreeEnergy for every window, and NumericalDifferentiableFunction.evaluateGradientOfFunction
    (jstacs :64-84) and AbstractFreeEnergyOptimizer.getWeightedLogLikelihoodSum (:95-114) over mean-field objects that share the
    model's net.  Restated: initParameters' hand-over of the new distribution to the tables (:49-53)."""
    _ns = None

    def __init__(self, net, windows, weights, position, eps=1e-4):
        if ReferenceObjective._ns is None:
            ReferenceEStep([1], 1, None, None, 0.5)              # (translates the Normalisation methods)
            ns = dict(ReferenceEStep._ns)
            with open(MATRIX_LINEARISATION_JAVA, encoding="latin-1") as f:
                ml = tokenize(f.read())
            exec(compile(_translate_method(ml, "lambda2pi", "lambda2pi", 3, [], []), MATRIX_LINEARISATION_JAVA, "exec"), ns)
            src = _translate_method(_jar_tokens(NUMDIFF_JAVA), "evaluateGradientOfFunction", "evaluateGradientOfFunction", 1,
                                    ["eps"], ["getDimensionOfScope", "evaluateFunction"])
            exec(compile(src, NUMDIFF_JAVA, "exec"), ns)
            path = REFERENCE_ROOT + "/main/java/src/optimizing/AbstractFreeEnergyOptimizer.java"
            with open(path, encoding="latin-1") as f:
                afe = tokenize(f.read())
            ssrc = _translate_method(afe, "getWeightedLogLikelihoodSum", "getWeightedLogLikelihoodSum", 0,
                                     ["vlfbn", "OPTIMIZE_MEANFIELDS", "weights"], [])
            exec(compile(ssrc, path, "exec"), ns)
            ReferenceObjective._ns, ReferenceObjective.gradient_source, ReferenceObjective.sum_source = ns, src, ssrc
        self.net, self.pos, self.eps = net, position, eps
        self.OPTIMIZE_MEANFIELDS = False                        # optimizing/AbstractFreeEnergyOptimizer.java:36
        self.weights = _JArray(float(w) for w in weights)
        # vlfbn: one MeanFieldForBayesNet per selected window over the model's one net, holding the mean fields the E-step's
        # scoring of that window left (models/PhyloBayesModel.java:139-147 collects them from the cache)
        self.base = ReferenceMeanField(net)
        self.base.bnh._myTreesArray = [_ObservedTree(t.nodes, self.base.leaf_nodes[i]) for i, t in enumerate(self.base.bnh._myTreesArray)]
        self.vlfbn = _JArray()
        for cols in windows:
            mf = _SharedMeanField(self.base, _MDSeq(cols))
            mf.initObservation()
            mf.init()
            mf.optimizeByNormalisation()
            self.vlfbn.append(mf)

    def getDimensionOfScope(self):
        return self.net.pi[self.pos].size

    def evaluateFunction(self, lam):
        pi = _JArray([0.0] * len(lam))
        self._ns["lambda2pi"](None, _JArray(lam), pi, 4)        # initParameters :49-53
        self.net.set_pi(self.pos, pi)
        self.net.build_tables()                                 # VirtualTree.initParametersFromGF: the tables of the changed position
        index = {n: i for i, n in enumerate(self.net.order_nodes)}
        for n, i in index.items():
            node = self.base.nodes[i]
            node.CPF = _CPF(node, self.net.tab[n].reshape(-1, 4), self.net.ind[n].reshape(-1, 4))
        return self._ns["getWeightedLogLikelihoodSum"](self)    # :95-114, executed

    def evaluateGradientOfFunction(self, lam):
        return list(self._ns["evaluateGradientOfFunction"](self, _JArray([float(x) for x in lam])))


# ---- the selector of the M-step (row a13) ---------------------------------------------------------------------------------------
SELECTOR_JAVA = REFERENCE_ROOT + "/main/java/src/io/SequenceSpecificDataSelector.java"


class _JList(list):
    """java.util.ArrayList / DoubleList as far as the translated methods use them: add / get / size"""

    def add(self, x):
        self.append(x)

    def get(self, i):
        return self[i]

    def size(self):
        return len(self)


class ReferenceSelector:
    """SequenceSpecificDataSelector for the windows of ONE alignment (io/SequenceSpecificDataSelector.java:37-99).  Executed
    from the reference: the walk of prepareForTraining (:82-96, its `for` statement) and Normalisation.sumNormalisation /
    normalisation (jstacs!/de/jstacs/utils/Normalisation.java:390-392,410-418,450-454).  Restated: the grouping by parent
    sequence and the ascending sort of transform() (AssociativeSort.quickSort; the order of equal weights is unspecified
    there -- a stable sort is used here and the tests avoid ties)."""
    _ns = None

    def __init__(self, weights):
        if ReferenceSelector._ns is None:
            ns = {"_jlog": _jlog, "_jexp": _jexp, "_NEG_INF": -math.inf, "_JArray": _JArray}
            nrm = _jar_tokens(NORMALISATION_JAVA)
            methods = ["sumNormalisation", "normalisation"]
            for py, java, n, sig in (("sum1", "sumNormalisation", 1, None), ("sum3", "sumNormalisation", 3, None),
                                     ("normd", "normalisation", 4, "double[] d, double v, double[] dest, int start")):
                exec(compile(_translate_method(nrm, py, java, n, [], methods, sig), NORMALISATION_JAVA, "exec"), ns)
            with open(SELECTOR_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            params, body = find_method(toks, "prepareForTraining", 2)
            start = next(i for i, t in enumerate(body) if t[1] == "for")            # behind the two collection declarations
            tr = Translator(["normWeightProfile", "weightProfile", "seqsPerParent"], [], params + ["seqs", "weights"])
            tr.emit(0, "def walk(self, t, q, seqs, weights):")
            tr.stmt(body, start, 1, [])
            src = "\n".join(tr.lines) + "\n"
            exec(compile(src, SELECTOR_JAVA, "exec"), ns)
            ReferenceSelector._ns, ReferenceSelector.source = ns, src
        order = sorted(range(len(weights)), key=lambda i: weights[i])                # transform(): ascending by weight
        self.seqsPerParent = _JArray([_JArray(order)])
        self.weightProfile = _JArray([_JArray(float(weights[i]) for i in order)])
        self.normWeightProfile = _JArray([_JArray(self.weightProfile[0])])
        self._ns["sum1"](self, self.normWeightProfile[0])                           # Normalisation.sumNormalisation(normWeightProfile[i])

    # static methods of Normalisation as the translated bodies call them
    def sumNormalisation(self, d, dest, start):
        return self._ns["sum3"](self, d, dest, start)

    def normalisation(self, d, v, dest, start):
        return self._ns["normd"](self, d, v, dest, start)

    def prepareForTraining(self, t, q):
        """-> (window indices, weights) in the order the reference adds them"""
        seqs, weights = _JList(), _JList()
        self._ns["walk"](self, t, q, seqs, weights)
        return list(seqs), list(weights)


def reference_log_sum(values):
    """Normalisation.getLogSum(double...) (jstacs!/de/jstacs/utils/Normalisation.java:43-45,63-93) -- what
    AbstractMixtureModel.getLogProbFor (jstacs :1158-1167) applies to the two strand terms of a StrandModel."""
    ns = reference_log_sum.__dict__.setdefault("ns", {})
    if not ns:
        ns.update({"_jlog": _jlog, "_jexp": _jexp, "_NEG_INF": -math.inf, "_JArray": _JArray, "_isinf": math.isinf})
        nrm = _jar_tokens(NORMALISATION_JAVA)
        exec(compile(_translate_method(nrm, "ls1", "getLogSum", 1, [], ["getLogSum"]), NORMALISATION_JAVA, "exec"), ns)
        exec(compile(_translate_method(nrm, "ls3", "getLogSum", 3, [], ["getLogSum"]), NORMALISATION_JAVA, "exec"), ns)

    class _Static:
        def getLogSum(self, *a):
            return ns["ls3"](self, *a) if len(a) == 3 else ns["ls1"](self, *a)
    return ns["ls1"](_Static(), _JArray(float(v) for v in values))


def reference_which_max(values):
    """Util.whichMax (util/Util.java:528-538): the arg-max rule of the site calls."""
    ns = reference_which_max.__dict__.setdefault("ns", {})
    if not ns:
        path = REFERENCE_ROOT + "/main/java/src/util/Util.java"
        with open(path, encoding="latin-1") as f:
            toks = tokenize(f.read())
        ns["_JArray"] = _JArray
        exec(compile(_translate_method(toks, "whichMax", "whichMax", 1, [], []), path, "exec"), ns)
    return ns["whichMax"](None, _JArray(float(v) for v in values))


# ---- PhyloBackground.getLogProbFor with the shared net (row a10, including ranges shorter than order + 1 columns) -------------
BACKGROUND_JAVA = REFERENCE_ROOT + "/main/java/src/models/PhyloBackground.java"


class _MDSeq:
    """MultiDimensionalDiscreteSequence as the path uses it: containerForPhyloBayes[column][leaf] (io/PhyloSample.java:106-136)"""

    def __init__(self, codes):
        self.containerForPhyloBayes = _JArray(_JArray(int(x) for x in col) for col in codes)
        self.learnedTopology = None

    def getLength(self):
        return len(self.containerForPhyloBayes)

    def getSubSequence(self, start, length):
        return _SubSeq(self, start, length)


class _SubSeq:
    def __init__(self, parent, start, length):
        self._parent, self.start, self._length = parent, start, length

    def getLength(self):
        return self._length

    def getParent(self):
        return self._parent


def _instanceof(x, name):
    return {"MultiDimensionalDiscreteSequence": _MDSeq, "SubSequence": _SubSeq}[name] is type(x)


class _SharedMeanField(ReferenceMeanField):
    """A MeanFieldForBayesNet of one (sub)sequence over the model's ONE net: the observation lives in the shared nodes
    (bayesNet/BayesNetNodeProperties.java:88-99), q and the hidden-node map are per object.  initObservation() is the
    reference's (:78-89), translated."""
    _init_obs = None

    def __init__(self, base, seq):
        if _SharedMeanField._init_obs is None:
            with open(REFERENCE_FILE, encoding="latin-1") as f:
                toks = tokenize(f.read())
            src = _translate_method(toks, "initObservation", "initObservation", 0, FIELDS, METHODS)
            ns = {"_instanceof": _instanceof, "_JArray": _JArray}
            exec(compile(src, REFERENCE_FILE, "exec"), ns)
            _SharedMeanField._init_obs, _SharedMeanField.init_observation_source = staticmethod(ns["initObservation"]), src
        self.net, self.nodes, self.bnh, self.leaf_nodes = base.net, base.nodes, base.bnh, base.leaf_nodes
        self.seq = seq
        self.parent = seq if _instanceof(seq, "MultiDimensionalDiscreteSequence") else seq.getParent()     # constructor :40-49
        self.q = self.node2hiddenNodeMap = self.ns = None
        self.hiddenNodes = self.free_energy_calls = 0

    def initObservation(self):
        return self._init_obs(self)


class _ObservedTree(_NodeList):
    def __init__(self, nodes, leaves):
        super().__init__(nodes)
        self.leaves = leaves

    def setObservation(self, obs):                              # bayesNet/VirtualTree.java:183-198, the leaves-only case
        assert len(obs) == len(self.leaves)
        for leaf, o in zip(self.leaves, obs):
            leaf.props.setObservation(o)

    def getNewickString(self):
        return ""


class ReferenceBackground:
    """PhyloBackground.getLogProbFor(MultiDimensionalDiscreteSequence, int, int) (models/PhyloBackground.java:241-305) executed
    from the reference over a net of order + 1 trees whose nodes every call shares -- so a range shorter than order + 1 columns
    leaves the trailing trees with whatever an earlier call observed there, exactly as the Java's call order produces it.
    Restated: getMFfromCache (models/AbstractPhyloModel.java:100-119: a new object per subsequence, initObservation, init), the
    empty score cache, check()."""
    _fn = None

    def __init__(self, net):
        if ReferenceBackground._fn is None:
            with open(BACKGROUND_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            fields = ["order", "bnh", "CALC_LOGLIKELIHOOD", "count", "_cachedScores"]
            methods = ["check", "isScoreCached", "getBNH", "reinitEdgelengths", "getMFfromCache", "prepareCache"]
            src = _translate_method(toks, "getLogProbFor", "getLogProbFor", 3, fields, methods,
                                    "MultiDimensionalDiscreteSequence sequence, int startpos, int endpos")
            ns = {"_JList": _JList, "_JArray": _JArray}
            exec(compile(src, BACKGROUND_JAVA, "exec"), ns)
            ReferenceBackground._fn, ReferenceBackground.source = staticmethod(ns["getLogProbFor"]), src
        self.base = ReferenceMeanField(net)
        self.base.bnh._myTreesArray = [_ObservedTree(t.nodes, self.base.leaf_nodes[i]) for i, t in enumerate(self.base.bnh._myTreesArray)]
        self.bnh, self.order = self.base.bnh, net.W - 1
        self.CALC_LOGLIKELIHOOD, self.count, self._cachedScores = False, 0, None
        self.sweeps = 0

    def check(self, *a):
        pass

    def isScoreCached(self, *a):
        return False

    def prepareCache(self, *a):
        pass

    def getBNH(self):
        bnh = self.bnh

        class _B:
            def getVirtualTree(self, i):
                return bnh._myTreesArray[i]
        return _B()

    def getMFfromCache(self, sequence):
        mf = _SharedMeanField(self.base, sequence)
        mf.initObservation()
        mf.init()
        return mf

    def getLogProbFor(self, sequence, startpos, endpos):
        return self._fn(self, sequence, startpos, endpos)


MOTIF_JAVA = REFERENCE_ROOT + "/main/java/src/models/PhyloBayesModel.java"


class ReferenceMotif(ReferenceBackground):
    """PhyloBayesModel.getLogProbFor(Sequence, int, int) (models/PhyloBayesModel.java:222-267) executed from the reference: the
    sub-sequence, the mean field from the (always missing) cache, initObservation, optimizeByNormalisation, -calcFreeEnergy()."""
    _motif_fn = None

    def __init__(self, net):
        if ReferenceMotif._motif_fn is None:
            with open(MOTIF_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            fields = ["bnh", "CALC_LOGLIKELIHOOD", "length"]
            methods = ["check", "getBNH", "reinitEdgelengths", "getMFfromCache", "getLength"]
            src = _translate_method(toks, "getLogProbFor", "getLogProbFor", 3, fields, methods, "Sequence sequence, int startpos, int endpos")
            ns = {"_instanceof": _instanceof, "_JArray": _JArray}
            exec(compile(src, MOTIF_JAVA, "exec"), ns)
            ReferenceMotif._motif_fn, ReferenceMotif.motif_source = staticmethod(ns["getLogProbFor"]), src
        self.base = ReferenceMeanField(net)
        self.base.bnh._myTreesArray = [_ObservedTree(t.nodes, self.base.leaf_nodes[i]) for i, t in enumerate(self.base.bnh._myTreesArray)]
        self.bnh, self.length, self.CALC_LOGLIKELIHOOD = self.base.bnh, net.W, False
        self.last = None

    def getLength(self):
        return self.length

    def getMFfromCache(self, sequence):
        self.last = super().getMFfromCache(sequence)
        return self.last

    def getLogProbFor(self, sequence, startpos, endpos):
        """-> (log-score, sweeps)"""
        v = self._motif_fn(self, sequence, startpos, endpos)
        return v, self.last.free_energy_calls - 2              # no-argument calcFreeEnergy: before the loop, per sweep, and the score


# ---- the non-phylogenetic motif model (row f-3): models/AlignmentBasedModel.java ------------------------------------------------
AB_JAVA = REFERENCE_ROOT + "/main/java/src/models/AlignmentBasedModel.java"


class _Row:
    """one species row of an alignment as Sequence.discreteVal sees it"""

    def __init__(self, codes):
        self.codes = [int(c) for c in codes]

    def discreteVal(self, i):
        return self.codes[i]


class ReferenceAlignmentBasedModel:
    """AlignmentBasedModel.getLnCondProb(sequence, posSeq, posMot) and isCompletylUnObserved (models/AlignmentBasedModel.java:
    187-228, 333-340) executed from the reference; cond[posMot] is the table of a motif position in the reference's index
    order, ACCESS_TABLE and pows as its constructor sizes them (:33,65-69)."""
    _ns = None

    def __init__(self, cond, order):
        if ReferenceAlignmentBasedModel._ns is None:
            with open(AB_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            fields = ["ACCESS_TABLE", "pows", "order", "length", "condProb", "IGNORE_GAPS_IN_LIKELIHOOD_CALC"]
            ns = {"_jlog": _jlog, "_JArray": _JArray, "_fill": _fill}
            src = _translate_method(toks, "getLnCondProb", "getLnCondProb", 3, fields, ["getLnCondProb"])
            usrc = _translate_method(toks, "isCompletylUnObserved", "isCompletylUnObserved", 3, [], [])
            exec(compile(src, AB_JAVA, "exec"), ns)
            exec(compile(usrc, AB_JAVA, "exec"), ns)

            class _Static:
                @staticmethod
                def isCompletylUnObserved(sequence, start, end):
                    return ns["isCompletylUnObserved"](None, sequence, start, end)
            ns["AlignmentBasedModel"] = _Static
            ReferenceAlignmentBasedModel._ns, ReferenceAlignmentBasedModel.source = ns, src
        self.order, self.length = order, len(cond)
        self.condProb = _JArray(_JArray(float(x) for x in row) for row in cond)
        self.pows = _JArray(4 ** i for i in range(order + 2))
        self.ACCESS_TABLE = _JArray([0] * sum(4 ** (i + 1) for i in range(order + 1)))
        self.IGNORE_GAPS_IN_LIKELIHOOD_CALC = False

    def getLnCondProb(self, row, pos_seq, pos_mot):
        return self._ns["getLnCondProb"](self, _Row(row), pos_seq, pos_mot)


AB_BG_JAVA = REFERENCE_ROOT + "/main/java/src/models/AlignmentBasedBGModel.java"


class ReferenceAlignmentBasedBackground:
    """AlignmentBasedBGModel.getLnCondProb(sequence, posSeq) (models/AlignmentBasedBGModel.java:173-222) executed from the
    reference -- including the list entries it keeps from before the last gap expansion; cond[min(posSeq, order)] is the table of
    that context length."""
    _ns = None

    def __init__(self, cond, order):
        if ReferenceAlignmentBasedBackground._ns is None:
            ReferenceAlignmentBasedModel([[0.25] * 4], 0)        # (translates isCompletylUnObserved)
            ns = dict(ReferenceAlignmentBasedModel._ns)
            ns.update({"_JList": _JList, "_copyOf": _copyOf})
            with open(AB_BG_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            fields = ["pows", "order", "condProb", "IGNORE_GAPS_IN_LIKELIHOOD_CALC"]
            src = _translate_method(toks, "getLnCondProbBG", "getLnCondProb", 2, fields, [], "Sequence<?> sequence, final int posSeq")
            exec(compile(src, AB_BG_JAVA, "exec"), ns)
            ReferenceAlignmentBasedBackground._ns, ReferenceAlignmentBasedBackground.source = ns, src
        self.order = order
        self.condProb = _JArray(_JArray(float(x) for x in row) for row in cond)
        self.pows = _JArray(4 ** i for i in range(order + 2))
        self.IGNORE_GAPS_IN_LIKELIHOOD_CALC = False

    def getLnCondProb(self, row, pos_seq):
        return self._ns["getLnCondProbBG"](self, _Row(row), pos_seq)


# ---- branch-length learning (row f-2): optimizing/branchlengths/EdgeOptimizer.java ----------------------------------------------
EDGE_JAVA = REFERENCE_ROOT + "/main/java/src/optimizing/branchlengths/EdgeOptimizer.java"


class _PhyloProps(_Props):
    def __init__(self, leaf, root):
        super().__init__(leaf)
        self._root = root

    def isPhyloRoot(self):
        return self._root


class ReferenceEdgeObjective(ReferenceObjective):
    """EdgeOptimizer.evaluateFunction (optimizing/branchlengths/EdgeOptimizer.java:44-74) executed from the reference: the
    logit-like map of the parameters to branch lengths, the hand-over to the nodes of the tree at `pos`, and
    getWeightedLogLikelihoodSum over fixed mean fields.  Restated: VirtualTree.initParametersFromGF (bayesNet/VirtualTree.java:
    82-96: the tables of a tree from its nodes' distances, values as the reinit statements give them)."""
    _edge = None

    def __init__(self, net, windows, weights, position, equal=False):
        super().__init__(net, windows, weights, position)
        if ReferenceEdgeObjective._edge is None:
            with open(EDGE_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            fields = ["MINIMUM_BRANCH_LENGTH", "bnh", "optimzeGlobally", "pos", "optimizeBranchlengthsEqual"]
            src = _translate_method(toks, "evaluateFunction", "evaluateFunction", 1, fields, ["getWeightedLogLikelihoodSum"])
            ns = dict(self._ns)
            ns["_isnan"] = math.isnan
            exec(compile(src, EDGE_JAVA, "exec"), ns)
            ReferenceEdgeObjective._edge, ReferenceEdgeObjective.edge_source = staticmethod(ns["evaluateFunction"]), src
        self.MINIMUM_BRANCH_LENGTH, self.optimzeGlobally, self.optimizeBranchlengthsEqual = 0.0005, False, equal
        self.bnh = self.base.bnh
        index = {n: i for i, n in enumerate(net.order_nodes)}
        owner = self

        for t, tree in enumerate(self.bnh._myTreesArray):
            for k, node in enumerate(tree.nodes):
                node.props.__class__ = _PhyloProps
                node.props._root = (k == 0)
                node.distance = net.dist[k]
                node.setDistanceToParent = (lambda v, node=node: setattr(node, "distance", float(v)))

            def rebuild(t=t, tree=tree):
                import numpy as np
                from oracle import np_restatement as npr
                for k, node in enumerate(tree.nodes):
                    if k == 0:
                        continue
                    tab = npr.f81alpha(owner.net.pi[t], node.distance)
                    node.CPF = _CPF(node, tab, np.array([owner.net.pi[t][l // 4] for l in range(tab.shape[0])]))
            tree.initParametersFromGF = rebuild
        self.bnh.getVirtualTree = lambda i: self.bnh._myTreesArray[i]

    def getWeightedLogLikelihoodSum(self):
        return self._ns["getWeightedLogLikelihoodSum"](self)

    def evaluateEdges(self, x):
        return self._edge(self, _JArray(float(v) for v in x))


# ---- the optimiser (row a16): jstacs Optimizer.quasiNewtonBFGS with its line search ---------------------------------------------
OPTIMIZER_JAVA = "de/jstacs/algorithms/optimization/Optimizer.java"


class ReferenceOptimizer:
    """Optimizer.quasiNewtonBFGS (jstacs!/de/jstacs/algorithms/optimization/Optimizer.java:1000-1080), findBracket (:122-146),
    brentsMethod (:202-303) and OneDimensionalSubFunction.evaluateFunction (:62-74) executed from the jar; the golden-section
    constants from their initialisers (:66-70).  Restated (handed in as objects): the termination condition and the start
    distance forecaster (phyfoo_b200.optimize.CombinedCondition / LimitedMedianStartDistance), DifferentiableFunction.
    findOneDimensionalMin -> OneDimensionalFunction.findMin (two lines: bracket, then Brent)."""
    _ns = None

    def __init__(self, fun, grad):
        if ReferenceOptimizer._ns is None:
            toks = _jar_tokens(OPTIMIZER_JAVA)
            ns = {"_JArray": _JArray, "_sqrt": math.sqrt, "_POS_INF": math.inf, "_NEG_INF": -math.inf}
            tr = Translator([], [], [])
            for name in ("limFib", "limFibPlus1", "limFibPlus2"):          # private static final double NAME = expr;
                at = next(q for q in range(len(toks) - 1) if toks[q] == ("id", name) and toks[q + 1][1] == "=")
                end = next(q for q in range(at, len(toks)) if toks[q][1] == ";")
                tr.locals.add(name)
                ns[name] = eval(tr.expr(toks[at + 2:end]), dict(ns))
            fields = ["limFib", "limFibPlus1", "limFibPlus2"]
            sigs = {"findBracket": (4, "OneDimensionalFunction f, double lower, double fLower, double startDistance"),
                    "brentsMethod": (6, "OneDimensionalFunction f, double a, double x, double fx, double b, double tol"),
                    "quasiNewtonBFGS": (7, None)}
            src = {}
            for name, (n, sig) in sigs.items():
                src[name] = _translate_method(toks, name, name, n, fields, ["checkStartDistance"], sig)
                exec(compile(src[name], OPTIMIZER_JAVA, "exec"), ns)
            sub = _jar_tokens("de/jstacs/algorithms/optimization/OneDimensionalSubFunction.java")
            src["along"] = _translate_method(sub, "along", "evaluateFunction", 1, ["f", "d", "current"], [])
            exec(compile(src["along"], OPTIMIZER_JAVA, "exec"), ns)
            ReferenceOptimizer._ns, ReferenceOptimizer.source = ns, src
        self.fun, self.grad, self.evaluations = fun, grad, 0
        self.limFib, self.limFibPlus1, self.limFibPlus2 = (self._ns[k] for k in ("limFib", "limFibPlus1", "limFibPlus2"))

    # DifferentiableFunction as the optimiser sees it
    def getDimensionOfScope(self):
        return self.n

    def evaluateFunction(self, x):
        self.evaluations += 1
        return float(self.fun(list(x)))

    def evaluateGradientOfFunction(self, x):
        self.evaluations += self.n + 1                           # NumericalDifferentiableFunction: f(x) and n perturbed points
        return _JArray(float(g) for g in self.grad(list(x)))

    def checkStartDistance(self, sd):
        if not sd > 0:
            raise ValueError("The start distance has to be positive.")

    def findOneDimensionalMin(self, x, d, alpha_0, f_alpha_0, lin_eps, start_distance):
        owner = self

        class _Sub:                                              # OneDimensionalSubFunction( this, x, d )
            f, current = owner, x

            def evaluateFunction(self, alpha):
                return owner._ns["along"](self, alpha)
        sub = _Sub()
        sub.d = d
        bracket = self._ns["findBracket"](self, sub, alpha_0, f_alpha_0, start_distance)        # OneDimensionalFunction.findMin :78-81
        return self._ns["brentsMethod"](self, sub, bracket[0], bracket[2], bracket[3], bracket[4], lin_eps)

    def minimise(self, x0, condition, forecaster, lin_eps=1e-3):
        """-> (x, iterations, function evaluations)"""
        self.n = len(x0)
        x = _JArray(float(v) for v in x0)

        class _Termination:
            def doNextIteration(self, i, f_last, f_cur, gradient, direction, alpha, t):
                return condition.do_next_iteration(i, f_last, f_cur, list(gradient), list(direction), alpha)

        class _Forecaster:
            def getNewStartDistance(self):
                return forecaster.get_new_start_distance()

            def setLastDistance(self, v):
                forecaster.set_last_distance(v)

        class _Time:
            def getElapsedTime(self):
                return 0.0
        it = self._ns["quasiNewtonBFGS"](self, self, x, _Termination(), lin_eps, _Forecaster(), None, _Time())
        return list(x), it, self.evaluations


TERMINATION_DIR = "de/jstacs/algorithms/optimization/termination/"
PREPARED_CONDITIONS_JAVA = REFERENCE_ROOT + "/main/java/src/optimizing/PreparedConditions.java"


def _sorted_inplace(a):
    a.sort()


def reference_condition(name="TC_MOTIF_SEPARATED_POSITIONS"):
    """The termination condition PhyFoo hands to the optimiser (optimizing/PreparedConditions.java:22-31), as an object with the
    do_next_iteration of phyfoo_b200.optimize.CombinedCondition: every doNextIteration (CombinedCondition,
    SmallDifferenceOfFunctionEvaluationsCondition, IterationCondition, SmallGradientConditon, SmallStepCondition) is executed
    from the jar; threshold and epsilons are read out of the constructor call in PreparedConditions.java."""
    with open(PREPARED_CONDITIONS_JAVA, encoding="latin-1") as f:
        text = re.sub(r"\s+", " ", f.read())
    call = re.search(name + r" = new CombinedCondition\((\d+),(.*?)\);", text)
    threshold = int(call.group(1))
    parts = re.findall(r"new (\w+)\(([-+.\deE]+)\)", call.group(2))
    field = {"SmallDifferenceOfFunctionEvaluationsCondition": "eps", "IterationCondition": "maxIter", "SmallGradientConditon": "eps",
             "SmallStepCondition": "epsilon"}

    class _Cond:
        pass
    conditions = _JArray()
    for cls, value in parts:
        toks = _jar_tokens(TERMINATION_DIR + cls + ".java")
        ns = {"_JArray": _JArray}
        exec(compile(_translate_method(toks, "doNextIteration", "doNextIteration", 7, [field[cls]], []), cls, "exec"), ns)
        c = _Cond()
        setattr(c, field[cls], int(value) if cls == "IterationCondition" else float(value))
        c.doNextIteration = (lambda *a, c=c, fn=ns["doNextIteration"]: fn(c, *a))
        conditions.append(c)
    toks = _jar_tokens(TERMINATION_DIR + "CombinedCondition.java")
    ns = {"_JArray": _JArray}
    src = _translate_method(toks, "doNextIteration", "doNextIteration", 7, ["condition", "threshold"], [])
    exec(compile(src, "CombinedCondition", "exec"), ns)
    combined = _Cond()
    combined.condition, combined.threshold, combined.source = conditions, threshold, src
    combined.do_next_iteration = lambda it, f_last, f_cur, grad, direction, alpha: ns["doNextIteration"](
        combined, it, f_last, f_cur, _JArray(grad), _JArray(direction), alpha, None)
    return combined


def reference_forecaster(slots=10, value=0.1):
    """LimitedMedianStartDistance (jstacs!/de/jstacs/algorithms/optimization/LimitedMedianStartDistance.java:64-84) executed from
    the jar: 0.667 times the median of the last `slots` step lengths, the memory initialised with `value` (:89-92)."""
    toks = _jar_tokens("de/jstacs/algorithms/optimization/LimitedMedianStartDistance.java")
    ns = {"_JArray": _JArray, "_arraycopy": lambda a, i, b, j, n: b.__setitem__(slice(j, j + n), a[i:i + n]), "_sort": _sorted_inplace}
    fields = ["initial", "memory", "sorted", "index"]
    get = _translate_method(toks, "getNewStartDistance", "getNewStartDistance", 0, fields, [], int_division=True)
    put = _translate_method(toks, "setLastDistance", "setLastDistance", 1, fields, [])
    for a, b in (("System . arraycopy", "_arraycopy"), ("Arrays . sort", "_sort")):
        get = get.replace(a, b)
    exec(compile(get, "LimitedMedianStartDistance", "exec"), ns)
    exec(compile(put, "LimitedMedianStartDistance", "exec"), ns)

    class _F:
        pass
    f = _F()
    f.initial, f.memory, f.sorted, f.index = value, _JArray([value] * slots), _JArray([0.0] * slots), 0      # constructor + reset()
    f.get_new_start_distance = lambda: ns["getNewStartDistance"](f)
    f.set_last_distance = lambda last: ns["setLastDistance"](f, last)
    f.source = get + put
    return f


STRAND_JAVA = "de/jstacs/models/mixture/StrandModel.java"
MIXTURE_JAVA = "de/jstacs/models/mixture/AbstractMixtureModel.java"


class ReferenceStrandEStep:
    """StrandModel.getNewWeights (jstacs!/de/jstacs/models/mixture/StrandModel.java:561-583) with AbstractMixtureModel.
    modifyWeights (EM branch, :1074-1078) and Normalisation.logSumNormalisation, executed from the jar: per window the strand
    posterior times the window's weight, the strand statistics w[2] and the log-likelihood L -- what the M-step of a mixture
    with a StrandModel component runs (row a11).  sample[0] holds the windows in pairs (forward, reverse complement)."""
    _ns = None

    def __init__(self, forward, reverse, forward_weight, hyper=(0.0, 0.0)):
        if ReferenceStrandEStep._ns is None:
            ReferenceEStep([1], 1, None, None, 0.5)
            ns = dict(ReferenceEStep._ns)
            fields, methods = ["model", "sample", "logWeights", "algorithm"], ["initWithPrior", "modifyWeights"]
            src = _translate_method(_jar_tokens(STRAND_JAVA), "strandNewWeights", "getNewWeights", 3, fields, methods)
            msrc = _translate_method(_jar_tokens(MIXTURE_JAVA), "modifyWeights", "modifyWeights", 1, fields, methods, case="EM")
            exec(compile(src, STRAND_JAVA, "exec"), ns)
            exec(compile(msrc, MIXTURE_JAVA, "exec"), ns)
            ReferenceStrandEStep._ns, ReferenceStrandEStep.source = ns, src + msrc
        scores = [x for pair in zip(forward, reverse) for x in pair]

        class _Model:
            def getLogProbFor(self, window):
                return float(scores[window])
        self.model, self.sample, self.hyper = [_Model()], [_Windows(len(scores))], hyper
        self.logWeights = _JArray([math.log(forward_weight), math.log(1.0 - forward_weight)])
        self.n = len(scores)

    def initWithPrior(self, w):
        w[0], w[1] = self.hyper

    def modifyWeights(self, w):
        return self._ns["modifyWeights"](self, w)

    def getNewWeights(self, data_weights=None):
        """-> (L, w[2], seqweights[0] in (forward, reverse) pairs)"""
        w, seqweights = _JArray([0.0, 0.0]), _JArray([_JArray([0.0] * self.n)])
        dw = None if data_weights is None else _JArray(float(x) for x in data_weights)
        L = self._ns["strandNewWeights"](self, dw, w, seqweights)
        return L, list(w), list(seqweights[0])


CLASSIFICATION_JAVA = REFERENCE_ROOT + "/main/java/src/classification/ClassificationUtil.java"


def reference_classification():
    """{name: callable}: determineThresholdsForSpecifitiesPerPos and determineSensitivityPerSeq of
    classification/ClassificationUtil.java (:244-278) executed from the reference (row f-1)."""
    with open(CLASSIFICATION_JAVA, encoding="latin-1") as f:
        toks = tokenize(f.read())
    ns = {"_JArray": _JArray, "_sort": _sorted_inplace}
    for name, n in (("determineThresholdsForSpecifitiesPerPos", 2), ("determineSensitivityPerSeq", 2)):
        exec(compile(_translate_method(toks, name, name, n, [], []), CLASSIFICATION_JAVA, "exec"), ns)
    wrap = lambda m: _JArray(_JArray(float(x) for x in row) for row in m)                       # noqa: E731
    return {"thresholds": lambda neg, spec: list(ns["determineThresholdsForSpecifitiesPerPos"](None, wrap(neg), _JArray(float(x) for x in spec))),
            "sensitivity": lambda thr, pos: ns["determineSensitivityPerSeq"](None, float(thr), wrap(pos))}


GLOBAL_EDGE_JAVA = REFERENCE_ROOT + "/main/java/src/optimizing/branchlengths/GlobalEdgeOptimizer.java"


class ReferenceGlobalEdgeObjective(ReferenceEdgeObjective):
    """GlobalEdgeOptimizer.evaluateFunction (optimizing/branchlengths/GlobalEdgeOptimizer.java:33-58) executed from the reference:
    one set of branch lengths for the trees of every motif position."""
    _global = None

    def __init__(self, net, windows, weights):
        super().__init__(net, windows, weights, 0)
        if ReferenceGlobalEdgeObjective._global is None:
            with open(GLOBAL_EDGE_JAVA, encoding="latin-1") as f:
                toks = tokenize(f.read())
            src = _translate_method(toks, "evaluateFunctionGlobal", "evaluateFunction", 1, ["bnh"], ["getWeightedLogLikelihoodSum"])
            ns = dict(self._ns)
            ns["_isnan"] = math.isnan

            class _EdgeOptimizer:
                MINIMUM_BRANCH_LENGTH = 0.0005                   # optimizing/branchlengths/EdgeOptimizer.java:23
            ns["EdgeOptimizer"] = _EdgeOptimizer
            exec(compile(src, GLOBAL_EDGE_JAVA, "exec"), ns)
            ReferenceGlobalEdgeObjective._global, ReferenceGlobalEdgeObjective.global_source = staticmethod(ns["evaluateFunctionGlobal"]), src

    def evaluateEdges(self, x):
        return self._global(self, _JArray(float(v) for v in x))
