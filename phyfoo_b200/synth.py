"""Deterministic synthetic alignments of the shapes BASELINE.json names.

Sequences are simulated root -> leaves under F81alpha exactly as the model defines it
(evolution/FS81alpha.java:76-92: a child keeps its parent's base with probability 1 - alpha, otherwise it
is redrawn from pi), background pi uniform like the bundled data, one planted motif site per alignment
with probability `zoops` at a uniform offset, optional iid leaf gaps.  Host-side numpy only; used by
bench.py and the tests to make inputs, never on the scoring path.
"""
import numpy as np

from .data import PhyloSample
from .newick import parse_newick


def star_newick(S, b=0.2):
    return "(" + ",".join(f"SPECIES_{i}:{b}" for i in range(S)) + ")"


def caterpillar_newick(S, b=0.2):
    s = f"(SPECIES_0:{b},SPECIES_1:{b})"
    for i in range(2, S):
        s = f"({s}:{b},SPECIES_{i}:{b})"
    return s


def random_binary_newick(S, seed, lo=0.05, hi=0.4):
    """Random join order; leaves are renamed so that SPECIES_i is the i-th leaf from the left."""
    rng = np.random.default_rng(seed)
    parts = [["L", None] for _ in range(S)]   # nested lists: ["L"] leaf or [left, right]
    nodes = [("L",) for _ in range(S)]
    while len(nodes) > 1:
        i, j = sorted(rng.choice(len(nodes), 2, replace=False))
        a, b = nodes[i], nodes[j]
        nodes = [n for k, n in enumerate(nodes) if k not in (i, j)] + [(a, b)]
    del parts
    counter = [0]

    def emit(n, top):
        if n == ("L",):
            s = f"SPECIES_{counter[0]}"
            counter[0] += 1
        else:
            s = "(" + emit(n[0], False) + "," + emit(n[1], False) + ")"
        return s if top else s + ":" + format(float(rng.uniform(lo, hi)), ".4f")

    return emit(nodes[0], True)


def random_pwm(W, order, seed):
    """Dirichlet(1,1,1,1) rows; position t of an order-1 motif has 4 context rows for t >= 1."""
    rng = np.random.default_rng(seed)
    out = []
    for t in range(W):
        rows = 4 if (order >= 1 and t > 0) else 1
        out.append(rng.dirichlet(np.ones(4), size=rows))
    return out


def _draw(rng, pi_rows, ctx, n):
    """n draws; row ctx[i] of pi_rows for draw i."""
    cdf = np.cumsum(pi_rows, axis=1)
    u = rng.random(n)
    c = cdf[ctx] if pi_rows.shape[0] > 1 else np.broadcast_to(cdf[0], (n, 4))
    return (u[:, None] > c[:, :3]).sum(axis=1).astype(np.uint8)


def simulate_columns(rng, tree, n, pi_rows=None, ctx=None):
    """n independent columns -> codes [n][S] under F81alpha on `tree` (parse_newick result)."""
    if pi_rows is None:
        pi_rows = np.full((1, 4), 0.25)
    if ctx is None:
        ctx = np.zeros(n, np.int64)
    N = tree.n_nodes
    vals = np.zeros((N, n), np.uint8)
    vals[0] = _draw(rng, pi_rows, ctx, n)
    for k in range(1, N):
        keep = rng.random(n) < (1.0 - tree.branch[k])
        vals[k] = np.where(keep, vals[tree.parent[k]], _draw(rng, pi_rows, ctx, n))
    leaves = np.argsort(np.where(tree.leaf_species >= 0, tree.leaf_species, 1 << 30))[: tree.n_species]
    return np.ascontiguousarray(vals[leaves].T), vals[0]


def make_sample(newick, n_aln, length, seed, pwm=None, zoops=1.0, gap_rate=0.0, length_jitter=0):
    """Returns (PhyloSample, planted offsets [-1 where none])."""
    rng = np.random.default_rng(seed)
    tree = parse_newick(newick)
    lens = np.full(n_aln, length, np.int64)
    if length_jitter:
        lens = lens + rng.integers(-length_jitter, length_jitter + 1, n_aln)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    codes, _ = simulate_columns(rng, tree, int(off[-1]))
    planted = np.full(n_aln, -1, np.int64)
    if pwm is not None:
        W = len(pwm)
        has = rng.random(n_aln) < zoops
        pos = (rng.random(n_aln) * (lens - W + 1)).astype(np.int64)
        idx = np.nonzero(has)[0]
        planted[idx] = pos[idx]
        ctx = np.zeros(idx.size, np.int64)
        for t in range(W):
            cols, root = simulate_columns(rng, tree, idx.size, np.asarray(pwm[t]), ctx)
            codes[off[idx] + pos[idx] + t] = cols
            ctx = root.astype(np.int64)
    if gap_rate > 0:
        gaps = rng.random(codes.shape) < gap_rate
        allgap = gaps.all(axis=1)
        gaps[allgap, 0] = False       # PhyloSample drops all-gap columns; keep lengths fixed instead
        codes = np.where(gaps, np.uint8(4), codes)
    return PhyloSample(codes, off), planted


# BASELINE.json configs (C2..C5).  W = 12 where BASELINE.json leaves it open.
CONFIGS = {
    "c2": dict(n_aln=2000, S=5, length=500, W=12, order=0, tree="caterpillar", seed=0x5EED0002),
    "c2star": dict(n_aln=2000, S=5, length=500, W=12, order=0, tree="star", seed=0x5EED0002),
    "c3": dict(n_aln=10000, S=12, length=1000, W=12, order=1, tree="random", seed=0x5EED0003),
    "c3o0": dict(n_aln=10000, S=12, length=1000, W=12, order=0, tree="random", seed=0x5EED0003),
    "c5": dict(n_aln=1000000, S=30, length=2000, W=12, order=0, tree="random", seed=0x5EED0005),
    # configs[3]: 3 motifs + homogeneous background, 100k alignments x 1 kb (sharded over 8 GPUs); one site of one of the
    # motifs in 80 % of the alignments
    "c4": dict(n_aln=100000, S=12, length=1000, W=12, order=0, tree="random", seed=0x5EED0004, motifs=3, zoops=0.8),
    # configs[0]: the reference's own bundled data (first 200 alignments of trees/data_tree1_1.0.fg, W = 10, generating
    # caterpillar tree); inputs from tests/golden/c1_tree1_first200.npz (scripts/make_c1_fixture.py)
    "c1": dict(n_aln=200, S=5, length=350, W=10, order=0, tree="caterpillar", seed=0, bundled=True),
}


def config_newick(cfg):
    if cfg["tree"] == "star":
        return star_newick(cfg["S"])
    if cfg["tree"] == "caterpillar":
        return caterpillar_newick(cfg["S"])
    return random_binary_newick(cfg["S"], cfg["seed"])


def make_config_c1(n_aln=None):
    """BASELINE configs[0]: the bundled alignments with their planted MOTIF_POS and the generating PWM / tree."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_tree1_first200.npz"))
    cfg = dict(CONFIGS["c1"])
    smp = PhyloSample(z["codes"], z["col_offsets"])
    planted = z["motif_pos"].astype(np.int64)
    if n_aln is not None and n_aln < smp.n_alignments:
        smp, planted = smp.slice(0, n_aln), planted[:n_aln]
    cfg["n_aln"] = smp.n_alignments
    return cfg, str(z["newick_cat"]), [p[None, :] for p in z["pwm"]], smp, planted


def make_config_multi(name, n_aln=None, shard=None):
    """A configuration with several motifs (c4).  shard = (rank, world): only this rank's block of n_aln / world alignments
    is generated, from a seed of its own (the whole set would be 1.2 GB of codes per process).  Returns
    (cfg, newick, [pwm_k], sample, planted offsets, planted motif index [-1 where none])."""
    cfg = dict(CONFIGS[name])
    if n_aln is not None:
        cfg["n_aln"] = int(n_aln)
    newick = config_newick(cfg)
    K = cfg["motifs"]
    pwms = [random_pwm(cfg["W"], cfg["order"], cfg["seed"] + 1 + k) for k in range(K)]
    rank, world = shard or (0, 1)
    lo, hi = cfg["n_aln"] * rank // world, cfg["n_aln"] * (rank + 1) // world
    rng = np.random.default_rng(cfg["seed"] + 1000 + rank)
    sample, _ = make_sample(newick, hi - lo, cfg["length"], cfg["seed"] + 2000 + rank)
    tree = parse_newick(newick)
    W = cfg["W"]
    n = hi - lo
    kind = np.where(rng.random(n) < cfg["zoops"], rng.integers(0, K, n), -1)
    pos = (rng.random(n) * (cfg["length"] - W + 1)).astype(np.int64)
    codes = sample.codes
    for k in range(K):
        idx = np.nonzero(kind == k)[0]
        for t in range(W):
            cols, _ = simulate_columns(rng, tree, idx.size, np.asarray(pwms[k][t]))
            codes[sample.col_offsets[idx] + pos[idx] + t] = cols
    planted = np.where(kind >= 0, pos, -1)
    return cfg, newick, pwms, PhyloSample(codes, sample.col_offsets), planted, kind


def make_config(name, n_aln=None, gap_rate=0.0):
    if name == "c1":
        return make_config_c1(n_aln)
    cfg = dict(CONFIGS[name])
    if n_aln is not None:
        cfg["n_aln"] = int(n_aln)
    newick = config_newick(cfg)
    pwm = random_pwm(cfg["W"], cfg["order"], cfg["seed"] + 1)
    sample, planted = make_sample(newick, cfg["n_aln"], cfg["length"], cfg["seed"] + 2, pwm=pwm, gap_rate=gap_rate)
    return cfg, newick, pwm, sample, planted
