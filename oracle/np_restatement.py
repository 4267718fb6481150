"""Second, independently written restatement of the same path (numpy tensor contractions).

TEST INFRASTRUCTURE ONLY (see oracle/phyfoo_oracle.cpp).  Its purpose is to pin the C++ oracle:
the reference ships no tests or golden vectors and no JVM exists here, so two restatements that
were written from the reference independently, in different styles, must agree to 1e-12.

The C++ oracle transliterates the reference's loops.  This file instead states the same
mathematics as a factor graph: every node n carries a tensor  T_n[val(par_k),...,val(par_0), val(n)]
(reference row index l = sum_p 4^p val(par_p): bayesNet/CPF.java:56-59,238-256) and the mean-field
update / free energy are expectations of log T_n under the product distribution q
(algorithm/MeanFieldForBayesNet.java:108-202 and :232-365).  Reference quirks kept on purpose:
 * an unobserved phylo leaf reads the "independent" table pi[ctx][b] (CPF.java:146-152,218-224);
 * in the UPDATE, the own-term of a hidden node sums over all values of an OBSERVED parent instead
   of clamping it (MeanFieldForBayesNet.java:117-135, TODO at :129); the free energy clamps (:305-311);
 * no max-shift in the normalisation (:192-201); stop rule |F_{k-1}-F_k| < 1e-5, sticky, 2 <= k <= 100.
Pure Python loops: small cases only.
"""
import math

import numpy as np


def parse_newick(s):
    """io/NewickToBayesNet.java:106-147: DFS pre-order numbering, parents before children."""
    parent, dist, label = [-1], [0.0], [""]
    cur, i, pending = 0, 0, None
    buf = ""

    def close(node):
        nonlocal buf, pending
        if buf:
            label[node] = buf
        if pending is not None:
            dist[node] = pending
        buf, pending = "", None

    want_dist = False
    while i < len(s):
        c = s[i]
        if c in "(),:;" or ord(c) <= 10:
            if c == "(":
                parent.append(cur); dist.append(0.0); label.append(""); cur = len(parent) - 1
            elif c == ",":
                close(cur)
                parent.append(parent[cur]); dist.append(0.0); label.append(""); cur = len(parent) - 1
            elif c == ")":
                close(cur); cur = parent[cur]
            elif c == ":":
                want_dist = True
            elif c == ";":
                break
            i += 1
            continue
        j = i
        while j < len(s) and s[j] not in "(),:;" and ord(s[j]) > 10:
            j += 1
        tok = s[i:j]
        if want_dist:
            pending, want_dist = float(tok), False
        else:
            buf += tok
        i = j
    n = len(parent)
    children = [[] for _ in range(n)]
    for k in range(1, n):
        children[parent[k]].append(k)
    leaves = [k for k in range(n) if not children[k]]
    return parent, dist, children, leaves, label


def f81alpha(pi_rows, alpha):
    """evolution/FS81alpha.java:76-92 -> table [a + 4*ctx][b]."""
    dim = pi_rows.shape[0]
    t = np.zeros((dim * 4, 4))
    for ctx in range(dim):
        for a in range(4):
            t[a + 4 * ctx] = alpha * pi_rows[ctx]
            t[a + 4 * ctx, a] += 1 - alpha
    return t


class Net:
    """W trees; horiz[t] = list of earlier trees whose node k is an extra parent of node k of tree t."""

    def __init__(self, newick, W, horiz):
        self.parent, self.dist, self.children, self.leaves, _ = parse_newick(newick)
        self.N, self.W, self.S = len(self.parent), W, len(self.leaves)
        self.par = {}
        for t in range(W):
            for k in range(self.N):
                p = [] if self.parent[k] < 0 else [(t, self.parent[k])]
                p += [(tt, k) for tt in horiz[t]]
                self.par[(t, k)] = p
        self.order_nodes = [(t, k) for t in range(W) for k in range(self.N)]
        self.chi = {n: [] for n in self.order_nodes}
        for t in range(W):                  # reference insertion order: tree children first (parse time), horizontal children when connected
            for k in range(self.N):
                for c in self.children[k]:
                    self.chi[(t, k)].append((t, c))
        for t in range(W):
            for tt in horiz[t]:
                for k in range(self.N):
                    self.chi[(tt, k)].append((t, k))
        self.pi = [np.full((4 ** len(horiz[t]), 4), 0.25) for t in range(W)]
        self.tab, self.ind = {}, {}
        self.obs = {}

    @staticmethod
    def motif(newick, W, order):
        return Net(newick, W, [[t - 1] if (order == 1 and t > 0) else [] for t in range(W)])

    @staticmethod
    def background(newick, order):
        return Net(newick, order + 1, [[t - j - 1 for j in range(order) if t - j > 0] for t in range(order + 1)])

    def set_pi(self, t, pi):
        self.pi[t] = np.asarray(pi, float).reshape(-1, 4)

    def build_tables(self):
        """bayesNet/VirtualTree.java:82-96 with F81alpha."""
        for t in range(self.W):
            for k in range(self.N):
                npar = len(self.par[(t, k)])
                if k == 0:
                    tab = self.pi[t].copy(); ind = self.pi[t].copy()
                else:
                    tab = f81alpha(self.pi[t], self.dist[k])
                    ind = np.array([self.pi[t][l // 4] for l in range(tab.shape[0])])
                shape = (4,) * npar + (4,)
                self.tab[(t, k)] = tab.reshape(shape)   # axes: par[npar-1], ..., par[0], value
                self.ind[(t, k)] = ind.reshape(shape)

    def observe(self, cols, revcomp=False):
        """cols [W][S] codes; >= 4 means unobserved (bayesNet/BayesNetNodeProperties.java:88-99)."""
        cols = np.asarray(cols)
        self.obs = {}
        for t in range(self.W):
            col = cols[self.W - 1 - t] if revcomp else cols[t]
            for s, k in enumerate(self.leaves):
                c = int(col[s])
                if revcomp and c < 4:
                    c = 3 - c
                if c < 4:
                    self.obs[(t, k)] = c

    def table_for(self, n):
        t, k = n
        if not self.children[k] and n not in self.obs:   # unobserved phylo leaf
            return self.ind[n]
        return self.tab[n]


def _expect_log(logT, weights):
    """Contract log-table axes (par[np-1],...,par[0],value) with one weight vector per axis
    (None keeps the axis).  Returns the remaining tensor."""
    out = logT
    keep = 0
    for w in weights:
        if w is None:
            keep += 1
        else:
            out = np.tensordot(out, w, axes=([keep], [0]))
    return out


def onehot(i):
    v = np.zeros(4); v[i] = 1.0
    return v


class MeanField:
    def __init__(self, net, n_trees_in_score):
        self.net, self.len = net, n_trees_in_score
        self.q = {n: np.full(4, 0.25) for n in net.order_nodes if n not in net.obs}
        self.sweeps = 0

    def _parent_weights(self, n, clamp_observed, skip=None):
        """weight vector per parent axis in tensor order (last parent first)."""
        ws = []
        for p in reversed(self.net.par[n]):
            if p == skip:
                ws.append(None)
            elif p in self.q:
                ws.append(self.q[p])
            else:
                ws.append(onehot(self.net.obs[p]) if clamp_observed else np.ones(4))
        return ws

    def free_energy(self, start=0, end=None):
        net = self.net
        end = self.len - 1 if end is None else end
        with np.errstate(divide="ignore", invalid="ignore"):
            part1 = 0.0
            for t in range(start, end + 1):
                for k in range(net.N):
                    n = (t, k)
                    if n in self.q:
                        qq = self.q[n]
                        part1 += sum(x * math.log(x) for x in qq if x > 0)
            part2 = 0.0
            for t in range(start, end + 1):
                for k in range(net.N):
                    n = (t, k)
                    tab = net.table_for(n)
                    logT = np.where(tab > 0, np.log(np.where(tab > 0, tab, 1.0)), -np.inf)
                    val_w = self.q[n] if n in self.q else onehot(net.obs[n])
                    if not net.par[n] and n in self.q:
                        part2 += sum(self.q[n][h] * logT[h] for h in range(4) if tab[h] > 0)
                        continue
                    e = _expect_log(logT, self._parent_weights(n, True) + [val_w])
                    part2 += float(e)
        return part1 - part2

    def sweep(self):
        net = self.net
        for n in net.order_nodes:
            if n not in self.q:
                continue
            s = np.zeros(4)
            logT = np.log(net.table_for(n))
            s += _expect_log(logT, self._parent_weights(n, False) + [None])
            for c in net.chi[n]:
                logTc = np.log(net.table_for(c))
                val_w = self.q[c] if c in self.q else onehot(net.obs[c])
                s += _expect_log(logTc, self._parent_weights(c, True, skip=n) + [val_w])
            e = np.exp(s)
            self.q[n] = e / e.sum()

    def optimize(self):
        k, conv = 0, False
        old = self.free_energy()
        while (k < 100 and not conv) or k < 2:
            self.sweep()
            new = self.free_energy()
            if abs(old - new) < 1e-5:
                conv = True
            old = new
            k += 1
        self.sweeps = k
        return k


def score_window(net, cols, revcomp=False):
    """models/PhyloBayesModel.java:222-267, mean-field mode."""
    net.observe(cols, revcomp)
    mf = MeanField(net, net.W)
    k = mf.optimize()
    return -mf.free_energy(), k


def exact_loglik_order0(net, cols):
    """Felsenstein pruning with gaps as missing data; what exact mode equals for order 0."""
    net.observe(cols)
    total = 0.0
    for t in range(net.W):
        part = np.ones((net.N, 4))
        for k in range(net.N - 1, 0, -1):
            tab = net.tab[(t, k)]
            if not net.children[k]:
                msg = tab[:, net.obs[(t, k)]] if (t, k) in net.obs else tab.sum(axis=1)
            else:
                msg = tab @ part[k]
            part[net.parent[k]] *= msg
        total += math.log(float(net.tab[(t, 0)].reshape(4) @ part[0]))
    return total


def star_closed_form(pi, alphas, col):
    """log sum_a pi_a prod_s [(1-alpha_s) delta(a,x_s) + alpha_s pi_{x_s}]  (hard part 5)."""
    tot = 0.0
    for a in range(4):
        p = pi[a]
        for s, x in enumerate(col):
            if x < 4:
                p *= (1 - alphas[s]) * (1.0 if a == x else 0.0) + alphas[s] * pi[x]
        tot += p
    return math.log(tot)


def bg_range_order0(net, codes, start, end):
    """models/PhyloBackground.java:241-305 for order 0: sum of independent per-column -F."""
    log_sum = 0.0
    for u in range(start, end + 1):
        net.observe(codes[u:u + 1])
        mf = MeanField(net, 1)
        mf.optimize()
        log_sum += mf.free_energy(0, 0)
    return -log_sum


def zoops_alignment(motif_scores, bg_cols, W, logw, weight=1.0):
    """jstacs SingleHiddenMotifMixture.getNewWeights (:520-564) for one alignment, background order 0.
    motif_scores[j] = motif log-score of window j; bg_cols[u] = background log-score of column u."""
    L = len(bg_cols)
    n = L - W + 1
    a = np.array([-math.log(L - W + 1) + motif_scores[j] - sum(bg_cols[j:j + W]) for j in range(n)])
    off = a.max()
    e = np.exp(a - off)
    z = off + math.log(e.sum())
    gam = e / e.sum()
    h = np.array([logw[0] + z, logw[1]])
    ho = h.max()
    he = np.exp(h - ho)
    lse = ho + math.log(he.sum())
    p = he / he.sum()
    ll = weight * (sum(bg_cols) + lse)
    return gam * p[0] * weight, p[0] * weight, p[1] * weight, ll, int(np.argmax(a))
