"""ctypes binding of the CPU oracle (oracle/libphyfoo_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under phyfoo_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

F81ALPHA, F81BETA, HKY, HKY_RATIO = 0, 1, 2, 3
MEANFIELD, EXACT_ELIMINATION, EXACT_PRUNING = 0, 1, 2
FG, BG = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "libphyfoo_oracle.so")
    src = os.path.join(_HERE, "phyfoo_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, u8p, ip = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_int)
        L.orc_last_error.restype = C.c_char_p
        L.orc_model_new.restype = C.c_void_p
        L.orc_model_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_double]
        L.orc_model_clone.restype = C.c_void_p
        L.orc_model_clone.argtypes = [C.c_void_p]
        L.orc_model_free.argtypes = [C.c_void_p]
        for f in ("orc_model_nodes", "orc_model_species", "orc_model_trees"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_model_pi_dim.argtypes = [C.c_void_p, C.c_int]
        L.orc_model_topology.argtypes = [C.c_void_p, ip, ip, dp]
        L.orc_model_leaf_label.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.orc_model_set_pi.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.orc_model_get_pi.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_model_set_branch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_model_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_model_cpf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp]
        L.orc_model_sweeps_total.restype = C.c_longlong
        L.orc_model_sweeps_total.argtypes = [C.c_void_p]
        L.orc_model_reset_stats.argtypes = [C.c_void_p]
        L.orc_score_window.restype = C.c_double
        L.orc_score_window.argtypes = [C.c_void_p, u8p, C.c_int, ip]
        L.orc_count_terms.argtypes = [C.c_void_p, u8p, C.c_int, dp]
        L.orc_score_windows.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, dp, ip]
        L.orc_exact_bruteforce.restype = C.c_double
        L.orc_exact_bruteforce.argtypes = [C.c_void_p, u8p]
        L.orc_bg_range.restype = C.c_double
        L.orc_bg_range.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, C.c_int]
        L.orc_bg_columns.argtypes = [C.c_void_p, u8p, C.c_int, dp]
        L.orc_estep.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int64), u8p, dp, dp, dp,
                                C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.POINTER(C.c_int32), u8p,
                                C.POINTER(C.c_longlong)]
        L.orc_ab_score_windows.argtypes = [dp, C.c_int, u8p, C.c_int, C.c_int, C.c_int, dp]
        L.orc_ab_bg_range.argtypes = [dp, u8p, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_ab_train.argtypes = [u8p, C.c_int, C.c_int, C.c_int, dp, C.c_double, dp]
        L.orc_select_group.argtypes = [dp, C.c_int, C.c_double, C.c_double, ip, dp]
        L.orc_lambda2pi.argtypes = [dp, dp, C.c_int]
        L.orc_fe_prepare.restype = C.c_void_p
        L.orc_fe_prepare.argtypes = [C.c_void_p, u8p, C.c_int, dp]
        L.orc_fe_free.argtypes = [C.c_void_p]
        L.orc_fe_sum.restype = C.c_double
        L.orc_fe_sum.argtypes = [C.c_void_p]
        L.orc_fe_eval.restype = C.c_double
        L.orc_fe_eval.argtypes = [C.c_void_p, C.c_int, dp, C.c_int]
        L.orc_fe_grad.argtypes = [C.c_void_p, C.c_int, dp, C.c_int, C.c_double, dp, dp]
        L.orc_fe_get_q.argtypes = [C.c_void_p, C.c_int, dp]
        L.orc_encode_alignment.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, u8p]
        L.orc_drop_allgap_columns.argtypes = [u8p, C.c_int, C.c_int]
        L.orc_pack_columns.argtypes = [u8p, C.c_int64, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _LIB = L
    return _LIB


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _err():
    return lib().orc_last_error().decode()


class Model:
    """One PhyloBayesModel (kind=FG) or PhyloBackground (kind=BG) of the oracle."""

    def __init__(self, kind, W, order, newick, evol=F81ALPHA, hky_ratio=0.0, _h=None):
        L = lib()
        self.h = _h or L.orc_model_new(kind, W, order, newick.encode(), evol, hky_ratio)
        if not self.h:
            raise ValueError(_err())
        self.kind, self.order = kind, order
        self.N = L.orc_model_nodes(self.h)
        self.S = L.orc_model_species(self.h)
        self.W = L.orc_model_trees(self.h)

    def clone(self):
        return Model(self.kind, self.W, self.order, "", _h=lib().orc_model_clone(self.h))

    def __del__(self):
        try:
            lib().orc_model_free(self.h)
        except Exception:
            pass

    def topology(self):
        parent = np.zeros(self.N, np.int32)
        leaf = np.zeros(self.N, np.int32)
        dist = np.zeros(self.N, np.float64)
        lib().orc_model_topology(self.h, parent.ctypes.data_as(C.POINTER(C.c_int)),
                                 leaf.ctypes.data_as(C.POINTER(C.c_int)), _dp(dist))
        return parent, leaf, dist

    def leaf_labels(self):
        out = []
        for s in range(self.S):
            buf = C.create_string_buffer(256)
            lib().orc_model_leaf_label(self.h, s, buf, 256)
            out.append(buf.value.decode())
        return out

    def pi_dim(self, t):
        return lib().orc_model_pi_dim(self.h, t)

    def set_pi(self, t, pi):
        pi = np.ascontiguousarray(pi, np.float64).ravel()
        if lib().orc_model_set_pi(self.h, t, _dp(pi), pi.size) != 0:
            raise ValueError(_err())

    def get_pi(self, t):
        out = np.zeros(self.pi_dim(t))
        lib().orc_model_get_pi(self.h, t, _dp(out))
        return out

    def set_branch(self, t, node, length):
        if lib().orc_model_set_branch(self.h, t, node, float(length)) != 0:
            raise ValueError(_err())

    def set_options(self, mode=MEANFIELD, allow_indep=True, log_in_loop=False):
        lib().orc_model_set_options(self.h, mode, int(allow_indep), int(log_in_loop))

    def cpf(self, t, k, indep=False):
        out = np.zeros(256 * 4)
        rows = lib().orc_model_cpf(self.h, t, k, int(indep), _dp(out))
        return out[:rows * 4].reshape(rows, 4).copy()

    def sweeps_total(self):
        return lib().orc_model_sweeps_total(self.h)

    def reset_stats(self):
        lib().orc_model_reset_stats(self.h)

    def score_window(self, cols, revcomp=False):
        cols = np.ascontiguousarray(cols, np.uint8)
        assert cols.shape == (self.W, self.S)
        k = C.c_int(0)
        v = lib().orc_score_window(self.h, _u8(cols), int(revcomp), C.byref(k))
        if np.isnan(v) and _err():
            raise RuntimeError(_err())
        return v, k.value

    def count_terms(self, cols, revcomp=False):
        """SURVEY.md section 8d term counter for one window: dict(upd_flops, fe_flops, upd_transc, fe_transc, sweeps)."""
        cols = np.ascontiguousarray(cols, np.uint8)
        out = np.zeros(5)
        if lib().orc_count_terms(self.h, _u8(cols), int(revcomp), _dp(out)) != 0:
            raise RuntimeError(_err())
        return dict(upd_flops=out[0], fe_flops=out[1], upd_transc=out[2], fe_transc=out[3], sweeps=int(out[4]))

    def score_windows(self, codes, revcomp=False):
        codes = np.ascontiguousarray(codes, np.uint8)
        Lc = codes.shape[0]
        n = Lc - self.W + 1
        out = np.zeros(max(n, 0))
        sw = np.zeros(max(n, 0), np.int32)
        if n > 0 and lib().orc_score_windows(self.h, _u8(codes), Lc, int(revcomp), _dp(out),
                                             sw.ctypes.data_as(C.POINTER(C.c_int))) != 0:
            raise RuntimeError(_err())
        return out, sw

    def exact_bruteforce(self, cols):
        cols = np.ascontiguousarray(cols, np.uint8)
        return lib().orc_exact_bruteforce(self.h, _u8(cols))

    def bg_range(self, codes, start, end):
        codes = np.ascontiguousarray(codes, np.uint8)
        return lib().orc_bg_range(self.h, _u8(codes), codes.shape[0], start, end)

    def bg_columns(self, codes):
        codes = np.ascontiguousarray(codes, np.uint8)
        out = np.zeros(codes.shape[0])
        if lib().orc_bg_columns(self.h, _u8(codes), codes.shape[0], _dp(out)) != 0:
            raise RuntimeError(_err())
        return out


def estep(fg, bg, col_offsets, codes, logw, strand_logw=None, aln_weights=None, faithful_cost=False,
          threads=1, want=("gamma", "fwd", "rev", "profile")):
    """SingleHiddenMotifMixture.getNewWeights over a whole sample. Returns a dict."""
    col_offsets = np.ascontiguousarray(col_offsets, np.int64)
    codes = np.ascontiguousarray(codes, np.uint8)
    n_aln = col_offsets.size - 1
    nwin = int(np.maximum(np.diff(col_offsets) - fg.W + 1, 0).sum())
    res = {k: (np.zeros(nwin) if k in want else None) for k in ("gamma", "fwd", "rev", "profile")}
    logw = np.ascontiguousarray(logw, np.float64)
    sl = None if strand_logw is None else np.ascontiguousarray(strand_logw, np.float64)
    aw = None if aln_weights is None else np.ascontiguousarray(aln_weights, np.float64)
    w = np.zeros(2)
    ll = C.c_double(0)
    arg_off = np.zeros(n_aln, np.int32)
    arg_strand = np.zeros(n_aln, np.uint8)
    sweeps = (C.c_longlong * 2)(0, 0)
    rc = lib().orc_estep(fg.h, bg.h, n_aln, col_offsets.ctypes.data_as(C.POINTER(C.c_int64)), _u8(codes),
                         _dp(logw), _dp(sl), _dp(aw), int(faithful_cost), int(threads), _dp(res["gamma"]),
                         _dp(res["fwd"]), _dp(res["rev"]), _dp(res["profile"]), _dp(w), C.byref(ll),
                         arg_off.ctypes.data_as(C.POINTER(C.c_int32)), _u8(arg_strand), sweeps)
    if rc != 0:
        raise RuntimeError(_err())
    res.update(w=w, ll=ll.value, argmax_off=arg_off, argmax_strand=arg_strand, sweeps=sweeps[0] + sweeps[1],
               sweeps_fg=sweeps[0], sweeps_bg=sweeps[1])
    return res


def select_group(w, t=0.8, q=0.0):
    w = np.ascontiguousarray(w, np.float64)
    idx = np.zeros(w.size, np.int32)
    wo = np.zeros(w.size)
    n = lib().orc_select_group(_dp(w), w.size, t, q, idx.ctypes.data_as(C.POINTER(C.c_int)), _dp(wo))
    return idx[:n].copy(), wo[:n].copy()


def lambda2pi(lam):
    lam = np.ascontiguousarray(lam, np.float64)
    out = np.zeros_like(lam)
    lib().orc_lambda2pi(_dp(lam), _dp(out), lam.size)
    return out


class FESet:
    """Fixed-q free-energy objective of the M-step over a list of windows (cols [n][W][S])."""

    def __init__(self, model, cols, weights):
        cols = np.ascontiguousarray(cols, np.uint8)
        weights = np.ascontiguousarray(weights, np.float64)
        self.model, self.n = model, cols.shape[0]
        self.h = lib().orc_fe_prepare(model.h, _u8(cols), self.n, _dp(weights))
        if not self.h:
            raise RuntimeError(_err())

    def __del__(self):
        try:
            lib().orc_fe_free(self.h)
        except Exception:
            pass

    def sum(self):
        return lib().orc_fe_sum(self.h)

    def eval(self, pos, lam):
        lam = np.ascontiguousarray(lam, np.float64)
        return lib().orc_fe_eval(self.h, pos, _dp(lam), lam.size)

    def grad(self, pos, lam, eps=1e-4):
        lam = np.array(lam, np.float64)
        g = np.zeros_like(lam)
        f0 = C.c_double(0)
        if lib().orc_fe_grad(self.h, pos, _dp(lam), lam.size, eps, C.byref(f0), _dp(g)) != 0:
            raise RuntimeError(_err())
        return f0.value, g

    def q(self, u):
        out = np.zeros(self.model.W * self.model.N * 4)
        k = lib().orc_fe_get_q(self.h, u, _dp(out))
        return out.reshape(self.model.W, self.model.N, 4), k


def encode_alignment(rows):
    S, Lc = len(rows), len(rows[0])
    arr = (C.c_char_p * S)(*[r.encode() if isinstance(r, str) else r for r in rows])
    codes = np.zeros((Lc, S), np.uint8)
    if lib().orc_encode_alignment(arr, S, Lc, _u8(codes)) != 0:
        raise ValueError(_err())
    return codes


def drop_allgap_columns(codes):
    codes = np.ascontiguousarray(codes, np.uint8).copy()
    n = lib().orc_drop_allgap_columns(_u8(codes), codes.shape[0], codes.shape[1])
    return codes[:n].copy()


def pack_columns(codes):
    codes = np.ascontiguousarray(codes, np.uint8)
    n, S = codes.shape
    bases = np.zeros(((S + 15) // 16, n), np.uint32)
    gaps = np.zeros(((S + 31) // 32, n), np.uint32)
    lib().orc_pack_columns(_u8(codes), n, S, bases.ctypes.data_as(C.POINTER(C.c_uint32)),
                           gaps.ctypes.data_as(C.POINTER(C.c_uint32)))
    return bases, gaps


# ---- non-phylogenetic alignment models, order 0 (models/AlignmentBasedModel.java, AlignmentBasedBGModel.java) ----
def ab_score_windows(condprob, codes, revcomp=False):
    condprob = np.ascontiguousarray(condprob, np.float64)
    codes = np.ascontiguousarray(codes, np.uint8)
    W = condprob.shape[0]
    out = np.zeros(max(codes.shape[0] - W + 1, 0))
    lib().orc_ab_score_windows(_dp(condprob), W, _u8(codes), codes.shape[0], codes.shape[1], int(revcomp), _dp(out))
    return out


def ab_bg_range(condprob4, codes, start, end):
    condprob4 = np.ascontiguousarray(condprob4, np.float64)
    codes = np.ascontiguousarray(codes, np.uint8)
    out = C.c_double()
    lib().orc_ab_bg_range(_dp(condprob4), _u8(codes), codes.shape[0], codes.shape[1], start, end, C.byref(out))
    return out.value


def ab_train(cols, weights, pseudo=1e-8):
    cols = np.ascontiguousarray(cols, np.uint8)
    weights = np.ascontiguousarray(weights, np.float64)
    n, W, S = cols.shape
    out = np.zeros((W, 4))
    lib().orc_ab_train(_u8(cols), n, W, S, _dp(weights), pseudo, _dp(out))
    return out


# ---- the same models for any order (test-sized inputs: plain loops that follow the Java statement by statement) ----
def _ab_pows(order):
    return [4 ** k for k in range(order + 2)]


def ab_motif_ln_cond_prob(cond, order, row, pos_seq, pos_mot):
    """AlignmentBasedModel.getLnCondProb(sequence, posSeq, posMot), models/AlignmentBasedModel.java:187-228.  cond[posMot] is the
    table of that motif position, row the codes of one species row (4 = unobserved)."""
    import math
    pows = _ab_pows(order)
    if all(row[i] == 4 for i in range(pos_seq - min(order, pos_mot), pos_seq + 1)):        # :189-191 isCompletylUnObserved
        return math.log(0.25)
    table = [0] * sum(pows[i] * 4 for i in range(order + 1))                                # ACCESS_TABLE, :33,65-69
    head = pre = 0
    k = 0
    while k <= order and k <= pos_mot:                                                      # :196
        if row[pos_seq - k] == 4:                                                           # :199-207
            old_pre = pre
            pre = head
            i = old_pre
            while i < pre or (pre == 0 and i == 0):
                for a in range(4):
                    table[head] = table[i] + a * pows[k]
                    head += 1
                i += 1
        else:                                                                               # :208-213
            head = 1 if head == 0 else head
            for i in range(pre, head):
                table[i] += int(row[pos_seq - k]) * pows[k]
        k += 1
    sum_prob = 0.0
    for i in range(pre, head):                                                              # :216-219
        sum_prob += cond[pos_mot][table[i]]
    return math.log(sum_prob / (head - pre))


def ab_score_windows_order(cond, order, codes, revcomp=False):
    """AlignmentBasedModel.getLogProbFor (:232-248) for every offset of one alignment; revcomp: the reverse complement of the
    window as jstacs StrandModel presents it (columns reversed, symbols complemented)."""
    codes = np.asarray(codes)
    W, (L, S) = len(cond), codes.shape
    out = np.zeros(max(L - W + 1, 0))
    for j in range(out.size):
        win = codes[j:j + W]
        if revcomp:
            win = np.where(win[::-1] == 4, 4, 3 - win[::-1])
        log_sum = 0.0
        for o in range(S):                                                                  # :241-243
            s = 0.0
            for i in range(W):                                                              # :170-175
                s += ab_motif_ln_cond_prob(cond, order, win[:, o], i, i)
            log_sum += s
        out[j] = log_sum
    return out


def ab_bg_ln_cond_prob(cond, order, row, pos_seq):
    """AlignmentBasedBGModel.getLnCondProb(sequence, posSeq), models/AlignmentBasedBGModel.java:172-222."""
    import math
    pows = _ab_pows(order)
    if all(row[i] == 4 for i in range(max(0, pos_seq - order), pos_seq + 1)):               # :174-176
        return math.log(0.25)
    observations = []
    pre_size = 0
    k = 0
    while k <= order and k <= pos_seq:                                                      # :180
        if row[pos_seq - k] == 4:                                                           # :183-199
            pre_size = len(observations)
            if len(observations) == 0:
                for a in range(4):
                    tmp = [0] * (order + 1)
                    tmp[k] = a
                    observations.append(tmp)
            else:
                for i in range(pre_size):                                                   # every earlier entry, :191
                    for a in range(4):
                        tmp = list(observations[i])
                        tmp[k] = a
                        observations.append(tmp)
        else:                                                                               # :200-208
            if len(observations) == 0:
                observations.append([0] * (order + 1))
            for obs in observations:
                obs[k] = int(row[pos_seq - k])
        k += 1
    sum_prob = 0.0
    size = len(observations) - pre_size
    for obs in observations[pre_size:]:                                                     # :211-219
        access = 0
        k = 0
        while k <= order and k <= pos_seq:
            access += pows[k] * obs[k]
            k += 1
        sum_prob += cond[min(pos_seq, order)][access] / size
    return math.log(sum_prob)


def ab_bg_range_order(cond, order, codes, start, end):
    """AlignmentBasedBGModel.getLogProbFor(sequence, start, end), :227-243."""
    codes = np.asarray(codes)
    if end < start:
        return 0.0
    log_sum = 0.0
    for l in range(start, end + 1):
        for o in range(codes.shape[1]):
            log_sum += ab_bg_ln_cond_prob(cond, order, codes[:, o], l)
    return log_sum


def _ab_normalise(counts):
    out = []
    for c in counts:                                                                        # :150-158 / :124-136
        c = np.asarray(c, np.float64).copy()
        for k in range(0, c.size, 4):
            c[k:k + 4] = c[k:k + 4] / (((c[k] + c[k + 1]) + c[k + 2]) + c[k + 3])
        out.append(c)
    return out


def ab_train_order(cols, weights, order, pseudo=1e-8):
    """AlignmentBasedModel.train, :82-168, on the selected windows cols[n][W][S] with their weights."""
    cols = np.asarray(cols)
    n, W, S = cols.shape
    pows = _ab_pows(order)
    counts = [np.full(pows[min(i, order)] * 4, pseudo) for i in range(W)]
    table = [0] * sum(pows[i] * 4 for i in range(order + 1))
    for w in range(n):
        for o in range(S):                                                                  # :120
            for l in range(W):
                for i in range(len(table)):
                    table[i] = 0
                head = pre = 0
                k = 0
                while k <= order and k <= l:                                                # :127
                    x = int(cols[w, l - k, o])
                    if x == 4:                                                              # :132-141
                        old_pre = pre
                        pre = head
                        i = old_pre
                        while i < pre or (pre == 0 and i == 0):
                            for a in range(4):
                                table[head] = table[i] + a * pows[k]
                                head += 1
                            i += 1
                    else:
                        head = 1 if head == 0 else head
                        for i in range(pre, head):
                            table[i] += x * pows[k]
                    k += 1
                for i in range(pre, head):                                                  # :149-151
                    counts[l][table[i]] += weights[w] / (head - pre)
    return _ab_normalise(counts)


def ab_bg_train_order(alignments, weights, order, pseudo=1e-2):
    """AlignmentBasedBGModel.train, :40-137; alignments: list of codes[L][S]."""
    pows = _ab_pows(order)
    counts = [np.full(pows[i] * 4, pseudo) for i in range(order + 1)]
    for w, codes in enumerate(alignments):
        codes = np.asarray(codes)
        for o in range(codes.shape[1]):
            for l in range(codes.shape[0]):
                observations = []
                pre_size = 0
                k = 0
                while k <= order and k <= l:
                    x = int(codes[l - k, o])
                    if x == 4:
                        old_pre = pre_size
                        pre_size = len(observations)
                        if len(observations) == 0:
                            for a in range(4):
                                tmp = [0] * (order + 1)
                                tmp[k] = a
                                observations.append(tmp)
                        else:
                            for i in range(old_pre, pre_size):                              # :91
                                for a in range(4):
                                    tmp = list(observations[i])
                                    tmp[k] = a
                                    observations.append(tmp)
                    else:
                        if len(observations) == 0:
                            observations.append([0] * (order + 1))
                        for obs in observations:
                            obs[k] = x
                    k += 1
                size = len(observations) - pre_size
                for obs in observations[pre_size:]:
                    h = 0
                    k = 0
                    while k <= order and k <= l:
                        h += pows[k] * obs[k]
                        k += 1
                    counts[min(l, order)][h] += (1.0 if weights is None else weights[w]) / size
    return _ab_normalise(counts)


# ---- ZOOPS mixture of K motif models (BASELINE.json configs[3]; no reference counterpart: jstacs SingleHiddenMotifMixture
# has one motif component, :716-718).  Specification = DESIGN.md "multi-motif mixture": the one-motif E-step
# (SingleHiddenMotifMixture.java:520-564, modify :585-596) per component, then ONE logSumNormalisation over the K + 1
# components.  Plain loops over the oracle's own window and column scores; K = 1 must reproduce estep() above.
def estep_multi(fgs, bg, col_offsets, codes, logw, aln_weights=None):
    col_offsets = np.ascontiguousarray(col_offsets, np.int64)
    codes = np.ascontiguousarray(codes, np.uint8)
    K, n_aln = len(fgs), col_offsets.size - 1
    logw = np.asarray(logw, np.float64)
    gam = [[] for _ in range(K)]
    w = np.zeros(K + 1)
    ll = 0.0
    arg_m, arg_o = np.zeros(n_aln, np.int32), np.zeros(n_aln, np.int32)
    for a in range(n_aln):
        aln = codes[col_offsets[a]:col_offsets[a + 1]]
        L = aln.shape[0]
        weight = 1.0 if aln_weights is None else float(aln_weights[a])
        bgc = bg.bg_columns(aln)                     # order 0: per-column log-scores; ranges are their sums in column order
        log_pseq = 0.0
        for u in range(L):
            log_pseq += bgc[u]
        Z, e_all, best_all = [], [], []
        for k, fg in enumerate(fgs):
            W = fg.W
            n = L - W + 1
            sc, _ = fg.score_windows(aln, False)
            aj = np.zeros(n)
            for j in range(n):
                bgw = 0.0
                for m in range(W):
                    bgw += bgc[j + m]
                aj[j] = -np.log(float(n)) + sc[j] - bgw
            off = aj.max()                                       # Normalisation.logSumNormalisation :352-375
            e = np.exp(aj - off)
            s = 0.0
            for v in e:
                s += v
            if s != 1.0:
                e = e / s
            Z.append(off + np.log(s)); e_all.append(e); best_all.append((off, int(np.argmax(aj))))
        help_ = np.array([logw[k] + Z[k] for k in range(K)] + [logw[K]])
        off = help_.max()
        h = np.exp(help_ - off)
        hs = 0.0
        for v in h:
            hs += v
        if hs != 1.0:
            h = h / hs
        ll += weight * (log_pseq + (off + np.log(hs)))
        best_v, best_k = -np.inf, 0
        for k in range(K):
            w[k] += weight * h[k]
            gam[k].append(e_all[k] * (h[k] * weight))
            v = logw[k] + best_all[k][0]
            if v > best_v:
                best_v, best_k = v, k
        w[K] += weight * h[K]
        arg_m[a], arg_o[a] = best_k, best_all[best_k][1]
    return {"gamma": [np.concatenate(g) for g in gam], "w": w, "ll": ll, "argmax_motif": arg_m, "argmax_off": arg_o}
