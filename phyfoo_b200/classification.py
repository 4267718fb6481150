"""Genome-scan / classification front-end: host-side mirror of classification/ClassificationUtil.java on top of the GPU
scores.  Two flavours per function: from a score matrix already on the host (numpy, the
reference's exact indexing) and straight from the device without shipping the scores (`*_device`, libphyfoo's
phyfoo_scan_hits / phyfoo_score_order_stats)."""
import ctypes as C

import numpy as np

from . import core
from .lib import lib, raise_for


def getScores(model, sample, strand=0):
    """double[][] getScores(model, sample) (:191-219): per alignment the scores of all offsets (empty if it is shorter
    than the motif)."""
    flat = model.getLogProbFor(sample, strand=strand)
    off = sample.device(model.ctx).window_offsets(model.getLength())
    return [flat[off[i]:off[i + 1]] for i in range(sample.n_alignments)]


def determineThresholdsForSpecifitiesPerPos(negativeScores, specifities):
    """:244-264: thresholds[s] = sorted(all negative scores)[(int)((n - 1) * specificity)]."""
    neg = np.sort(np.concatenate([np.asarray(x, np.float64) for x in negativeScores]))
    return np.array([neg[int((neg.size - 1) * s)] for s in specifities])


def determineThresholdForSpecifity(specifity, scores):
    """:161-180 on a score matrix: sorted(all scores)[(int)(n * specificity)]."""
    flat = np.sort(np.concatenate([np.asarray(x, np.float64) for x in scores]))
    return float(flat[int(flat.size * specifity)])


def determineSensitivityPerSeq(threshold, positiveScores):
    """:266-278: fraction of alignments with at least one score > threshold."""
    return sum(1 for x in positiveScores if (np.asarray(x) > threshold).any()) / len(positiveScores)


def getSensitivitiesPerSeqForSpecifitiesPerPos(positiveScores, negativeScores, specifities):
    """:231-242."""
    thr = determineThresholdsForSpecifitiesPerPos(negativeScores, specifities)
    return np.array([determineSensitivityPerSeq(t, positiveScores) for t in thr])


def determinePositionSet(threshold, positiveScores):
    """:287-298: every (alignment, offset) whose score exceeds the threshold."""
    return {(i, int(u)) for i, x in enumerate(positiveScores) for u in np.nonzero(np.asarray(x) > threshold)[0]}


def calculateAUC(x, y):
    """:57-69 trapezoid rule starting at the origin."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    if x.size != y.size:
        raise ValueError("x and y must have the same length.")
    area = (x[0] - 0) * ((y[0] - 0) / 2 + 0)
    for i in range(1, x.size):
        area += (x[i] - x[i - 1]) * ((y[i] - y[i - 1]) / 2 + y[i - 1])
    return float(area)


def getDifferences(fgScores, bgScores):
    """:82-90 on per-alignment log-likelihoods."""
    return np.asarray(fgScores, np.float64) - np.asarray(bgScores, np.float64)


def getDynamicClassificationRate(fgDiffs, bgDiffs, thresh=None):
    """:96-121 (given threshold) and :127-159 (best threshold among the foreground differences)."""
    fg, bg = np.sort(np.asarray(fgDiffs, np.float64)), np.sort(np.asarray(bgDiffs, np.float64))
    if thresh is not None:
        tp, fp = (fg > thresh).sum(), (bg > thresh).sum()
        return float(tp + (bg.size - fp)) / (fg.size + bg.size)
    best = 0.0
    for i in range(fg.size - 1, -1, -1):
        tp, fn = fg.size - i, i
        fp = (bg > fg[i]).sum()
        best = max(best, float(tp + bg.size - fp) / (tp + bg.size + fn))
    return best


# ---- device-side flavours: the scores never leave the GPU -----------------------------------------------------------
def scan_hits_device(model, sample, threshold, strand=0):
    """Per alignment: (#offsets with score > threshold, first such offset or -1, maximal score, its first offset)."""
    ctx = model.ctx
    ds = sample.device(ctx)
    n = ds.n_aln
    hits, first, arg = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    best = np.zeros(n)
    raise_for(lib().phyfoo_scan_hits(ctx.h, ds.h, model.device_model().h, int(strand), float(threshold),
                                     core._p(hits, C.c_int32), core._p(first, C.c_int32), core._p(best, C.c_double),
                                     core._p(arg, C.c_int32)), ctx.h)
    return {"n_hits": hits, "first_hit": first, "max_score": best, "arg_max": arg}


def determineSensitivityPerSeq_device(threshold, model, sample, strand=0):
    """determineSensitivity(threshold, model, sample) (:300-303) without shipping the scores."""
    h = scan_hits_device(model, sample, threshold, strand)
    return float((h["n_hits"] > 0).sum()) / sample.n_alignments


def determineThresholdsForSpecifitiesPerPos_device(model, negativeSample, specifities, strand=0, rule="per_pos"):
    """rule 'per_pos': rank (int)((n - 1) * s) (:244-264); rule 'single': rank (int)(n * s) (:161-180)."""
    ctx = model.ctx
    ds = negativeSample.device(ctx)
    n = C.c_int64(0)
    raise_for(lib().phyfoo_score_order_stats(ctx.h, ds.h, model.device_model().h, int(strand), 0, None, None, C.byref(n)), ctx.h)
    ranks = np.array([int((n.value - 1) * s) if rule == "per_pos" else int(n.value * s) for s in specifities], np.int64)
    out = np.zeros(ranks.size)
    raise_for(lib().phyfoo_score_order_stats(ctx.h, ds.h, model.device_model().h, int(strand), ranks.size,
                                             core._p(ranks, C.c_int64), core._p(out, C.c_double), C.byref(n)), ctx.h)
    return out
