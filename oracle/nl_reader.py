"""Independent (pure-Python) AMPL .nl reader used ONLY as a checker for the product's C++ reader
(bonmin_b200/csrc/bnlp_nl.cpp).  TEST INFRASTRUCTURE ONLY.  Evaluates expression trees directly."""
from __future__ import annotations

import math

import numpy as np

INF = 1e19
_UN = {15: abs, 16: lambda a: -a, 37: math.tanh, 38: math.tan, 39: math.sqrt, 40: math.sinh, 41: math.sin,
       42: math.log10, 43: math.log, 44: math.exp, 45: math.cosh, 46: math.cos, 49: math.atan, 51: math.asin, 53: math.acos,
       76: lambda a: a * a}
_BIN = {0: lambda a, b: a + b, 1: lambda a, b: a - b, 2: lambda a, b: a * b, 3: lambda a, b: a / b,
        5: lambda a, b: a ** b, 75: lambda a, b: a ** b, 77: lambda a, b: a ** b}


class NlModel:
    def __init__(self, text: str):
        lines = []
        for ln in text.splitlines():
            ln = ln.split("#")[0].strip()
            if ln:
                lines.append(ln)
        self._l, self._p = lines, 0
        assert self._next()[0] == "g"
        h = [int(float(t)) for t in self._next().split()]
        self.n, self.m = h[0], h[1]
        self._next(); self._next()
        h5 = [int(t) for t in self._next().split()]
        self.nlvc, self.nlvo, self.nlvb = h5[:3]
        self._next()
        self.nbv, self.niv, self.nlvbi, self.nlvci, self.nlvoi = [int(t) for t in self._next().split()][:5]
        self._next(); self._next(); self._next()
        n, m = self.n, self.m
        self.xl = np.full(n, -INF); self.xu = np.full(n, INF)
        self.gl = np.full(m, -INF); self.gu = np.full(m, INF)
        self.x0 = np.zeros(n)
        self.expr = {k: ("n", 0.0) for k in range(m + 1)}
        self.lin = np.zeros((m + 1, n))
        self.sense = 0
        self.suffixes = {}
        self.defs = []      # defined variables (V segments): (linear terms, expression); index n + i
        while self._p < len(self._l):
            s = self._next()
            k = s[0]
            if k == "V":
                idx, nlin, _ = [int(t) for t in s[1:].split()]
                assert idx == n + len(self.defs)
                lin = []
                for _ in range(nlin):
                    v, a = self._next().split()
                    lin.append((int(v), float(a)))
                self.defs.append((lin, self._expr()))
            elif k == "C":
                self.expr[int(s[1:]) + 1] = self._expr()
            elif k == "O":
                self.sense = int(s[1:].split()[1])
                self.expr[0] = self._expr()
            elif k in "rb":
                for i in range(m if k == "r" else n):
                    t = self._next().split()
                    code = int(t[0])
                    lo, up = -INF, INF
                    if code == 0:
                        lo, up = float(t[1]), float(t[2])
                    elif code == 1:
                        up = float(t[1])
                    elif code == 2:
                        lo = float(t[1])
                    elif code == 4:
                        lo = up = float(t[1])
                    if k == "r":
                        self.gl[i], self.gu[i] = lo, up
                    else:
                        self.xl[i], self.xu[i] = lo, up
            elif k == "k":
                for _ in range(int(s[1:])):
                    self._next()
            elif k in "JG":
                idx, cnt = [int(t) for t in s[1:].split()]
                fn = idx + 1 if k == "J" else 0
                for _ in range(cnt):
                    v, a = self._next().split()
                    self.lin[fn, int(v)] += float(a)
            elif k == "x":
                for _ in range(int(s[1:])):
                    v, a = self._next().split()
                    self.x0[int(v)] = float(a)
            elif k == "S":      # suffix: "S<kind> <count> <name>", kind & 3 = 0 variables / 1 constraints / 2 objectives / 3 problem, kind & 4 = real
                kind, cnt, name = s[1:].split()
                tab = self.suffixes.setdefault((int(kind) & 3, name), {})
                for _ in range(int(cnt)):
                    v, a = self._next().split()
                    tab[int(v)] = float(a) if int(kind) & 4 else int(float(a))
            else:
                raise ValueError(s)

    def branching_info(self):
        """AmplTMINLP::read_priorities (src/Interfaces/Ampl/BonAmplTMINLP.cpp:215-252): variable suffixes priority (stored
        as 9999 - value: Cbc branches on the LOWEST number first), direction, upPseudocost / downPseudocost (one given:
        used for both).  Entries are None when the suffix is absent; unlisted variables have suffix value 0."""
        def dense(name, dtype):
            tab = self.suffixes.get((0, name))
            if tab is None:
                return None
            a = np.zeros(self.n, dtype)
            for i, v in tab.items():
                a[i] = v
            return a
        pri = dense("priority", np.int64)
        up, dn = dense("upPseudocost", float), dense("downPseudocost", float)
        if up is not None and dn is None:
            dn = up
        elif dn is not None and up is None:
            up = dn
        return dict(priorities=None if pri is None else 9999 - pri, directions=dense("direction", np.int64), up=up, down=dn)

    def sos_info(self):
        """AmplTMINLP::read_sos (BonAmplTMINLP.cpp:254-296) from the .sosno / .ref variable suffixes: sets by increasing set
        number, type 1 for a positive number and 2 for a negative one, members by increasing reference weight, weights made
        strictly increasing (sos_kludge.cpp:29-45), priority 9999 - the largest member priority."""
        no = self.suffixes.get((0, "sosno"))
        if not no:
            return None
        ref = self.suffixes.get((0, "ref")) or {}
        pri = self.suffixes.get((0, "priority")) or {}
        mem = sorted((int(s), float(ref.get(j, 0.0)) if ref else float(j), j) for j, s in no.items() if int(s) != 0)
        types, prios, starts, idx, wts = [], [], [], [], []
        for e, (s, w, j) in enumerate(mem):
            if e == 0 or s != mem[e - 1][0]:
                starts.append(e)
                types.append(1 if s > 0 else 2)
                prios.append(0)
            idx.append(j)
            wts.append(w)
            prios[-1] = max(prios[-1], int(pri.get(j, 0)))
        starts.append(len(mem))
        for k in range(len(types)):
            t = wts[starts[k]]
            for e in range(starts[k] + 1, starts[k + 1]):
                t += 1e-10
                if wts[e] <= t:
                    wts[e] = t + 1e-10
                t = wts[e]
        return dict(types=np.array(types), priorities=9999 - np.array(prios), starts=np.array(starts), indices=np.array(idx),
                    weights=np.array(wts))

    def _next(self):
        s = self._l[self._p]
        self._p += 1
        return s

    def _expr(self):
        s = self._next()
        if s[0] == "n":
            return ("n", float(s[1:]))
        if s[0] == "v":
            return ("v", int(s[1:]))
        op = int(s[1:])
        if op == 54:
            return ("sum", [self._expr() for _ in range(int(self._next()))])
        if op in _UN:
            return ("u", op, self._expr())
        if op in _BIN:
            a = self._expr()
            return ("b", op, a, self._expr())
        raise ValueError(f"opcode {op}")

    def _ev(self, e, x):
        t = e[0]
        if t == "n":
            return e[1]
        if t == "v":
            if e[1] >= self.n:
                lin, ex = self.defs[e[1] - self.n]
                return sum(a * self._ev(("v", v), x) for v, a in lin) + self._ev(ex, x)
            return x[e[1]]
        if t == "sum":
            return sum(self._ev(k, x) for k in e[1])
        if t == "u":
            return _UN[e[1]](self._ev(e[2], x))
        return _BIN[e[1]](self._ev(e[2], x), self._ev(e[3], x))

    def values(self, x):
        """[f, g_1..g_m] at x (minimisation form: a maximised objective is negated)."""
        v = np.array([self._ev(self.expr[k], x) + self.lin[k] @ x for k in range(self.m + 1)])
        if self.sense == 1:
            v[0] = -v[0]
        return v

    def var_types(self):
        """BonAmplTMINLP.cpp:473-564."""
        n = self.n
        vt = np.zeros(n, int)
        end = self.nlvb - self.nlvbi
        start, end = end, end + self.nlvbi
        vt[start:end] = 2
        start, end = end, max(end, end + self.nlvc - self.nlvci - self.nlvb)
        start, end = end, end + self.nlvci
        vt[start:end] = 2
        start, end = end, max(end, end + self.nlvo - max(self.nlvb, self.nlvc) - self.nlvoi)
        start, end = end, end + self.nlvoi
        vt[start:end] = 2
        start = n - self.nbv - self.niv
        vt[start:start + self.nbv] = 1
        vt[start + self.nbv:n] = 2
        return vt
