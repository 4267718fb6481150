"""Alignments of orthologous promoters as a column tensor (host side) + the device-resident packed copy.

Mirrors io/PhyloSample.java (reference): FASTA records annotated `>gene=<id>; species=<name>; ...;`
(io/SampleUtil.java:58-66) are grouped by gene into one alignment each (:75-138), species rows are ordered
by the leaf order of the Newick string (util/Util.java:39-53), columns that are gaps in every species are
dropped (:196-223), and symbols are coded A,C,G,T,- -> 0..4 case-insensitively
(jstacs UnobservableDNAAlphabet.java:116).  The reference iterates genes in Java HashSet order; here genes keep
the order of their first record.
"""
import re

import numpy as np

from . import core
from .newick import species_order

_CODE = np.full(256, 255, np.uint8)
for _i, _ch in enumerate("ACGT"):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i
_CODE[ord("-")] = 4


class WrongAlphabetException(ValueError):
    pass


def encode(seq: str) -> np.ndarray:
    raw = np.frombuffer(seq.encode("ascii"), np.uint8)
    out = _CODE[raw]
    if (out == 255).any():
        bad = chr(int(raw[np.argmax(out == 255)]))
        raise WrongAlphabetException(f"symbol {bad!r} is not in the alphabet {{A,C,G,T,-}}")
    return out


def parse_annotation(header: str) -> dict:
    """SplitSequenceAnnotationParser("=", ";"): `key=value; key=value;`"""
    out = {}
    for part in header.lstrip(">").split(";"):
        part = part.strip()
        if "=" in part:
            k, v = part.split("=", 1)
            out[k.strip()] = v.strip()
    return out


def read_fasta(path):
    records = []
    header, chunks = None, []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if header is not None:
                    records.append((header, "".join(chunks)))
                header, chunks = line, []
            elif line:
                chunks.append(line)
    if header is not None:
        records.append((header, "".join(chunks)))
    return records


_BRANCH = re.compile(r":([0-9]*\.[0-9]*)")          # the pattern of io/PhyloSample.java:283 and bayesNet/VirtualTree.java:240


def combineBranchLengths(newick1: str, newick2: str, weight: float) -> str:
    """io/PhyloSample.java:272-310: newick1 with every branch length replaced by weight * its own + (1 - weight) * the
    length at the same place of newick2 (lengths are matched by their order of appearance), printed like
    String.valueOf(double)."""
    from .xmlio import jdouble
    d1 = [float(m.group(1)) for m in _BRANCH.finditer(newick1)]
    d2 = [float(m.group(1)) for m in _BRANCH.finditer(newick2)]
    if len(d2) < len(d1):
        raise IndexError("the second tree has fewer branch lengths than the first")       # DoubleList.get out of range
    out, last = [], 0
    for s, m in enumerate(_BRANCH.finditer(newick1)):
        out.append(newick1[last:m.start(1)])
        out.append(jdouble(d1[s] * weight + d2[s] * (1 - weight)))
        last = m.end(1)
    out.append(newick1[last:])
    return "".join(out)


class PhyloSample:
    """A sample of multiple alignments: codes [sum L_i][S] (uint8, 0..4) + col_offsets [n+1].  `learnedTopology` (None, or one
    Newick string or None per alignment; set by training.learnTopologies = PhyloSample.train of the reference,
    io/PhyloSample.java:243-270) makes the models score an alignment under its own branch lengths."""

    combineBranchLengths = staticmethod(combineBranchLengths)

    def __init__(self, codes, col_offsets, species=None, genes=None, annotations=None):
        self.codes = np.ascontiguousarray(codes, np.uint8)
        self.col_offsets = np.ascontiguousarray(col_offsets, np.int64)
        if self.codes.ndim != 2 or self.col_offsets[0] != 0 or self.col_offsets[-1] != self.codes.shape[0]:
            raise ValueError("codes must be [columns][species] and col_offsets its alignment boundaries")
        self.species = list(species) if species is not None else None
        self.genes = list(genes) if genes is not None else None
        self.annotations = annotations
        self.learnedTopology = None
        self._device = {}

    # ---- constructors -------------------------------------------------------------------------
    @classmethod
    def from_alignments(cls, alignments, **kw):
        """alignments: list of [L_i][S] code arrays."""
        lens = [a.shape[0] for a in alignments]
        codes = np.concatenate(alignments, axis=0) if alignments else np.zeros((0, 1), np.uint8)
        return cls(codes, np.concatenate([[0], np.cumsum(lens)]), **kw)

    @classmethod
    def from_fasta(cls, path, newick, species_ident="species", gene_ident="gene", limit=None):
        order = species_order(newick)
        index = {name: i for i, name in enumerate(order)}
        groups, gene_order, annots = {}, [], {}
        for header, seq in read_fasta(path):
            ann = parse_annotation(header)
            gene, sp = ann.get(gene_ident), ann.get(species_ident)
            if gene is None or sp is None:
                raise ValueError(f"record without {gene_ident}=/{species_ident}= annotation: {header}")
            if gene not in groups:
                groups[gene] = {}
                gene_order.append(gene)
                annots[gene] = ann
            groups[gene][sp] = seq
        seen = set(sp for g in groups.values() for sp in g)
        for name in order:
            if name not in seen:
                raise IOError(f"A species-name ({name}) exisiting in the tree does not exist in the data-set.")
        alignments, genes = [], []
        for gene in gene_order[:limit]:
            rows = groups[gene]
            if any(name not in rows for name in order):
                raise ValueError(f"gene {gene}: missing species")
            mat = np.stack([encode(rows[name]) for name in order], axis=1)      # [L][S]
            keep = (mat == 4).sum(axis=1) < mat.shape[1]                          # PhyloSample.java:196-223
            alignments.append(np.ascontiguousarray(mat[keep]))
            genes.append(gene)
        return cls.from_alignments(alignments, species=order, genes=genes, annotations=[annots[g] for g in genes])

    # ---- accessors ----------------------------------------------------------------------------
    @property
    def n_alignments(self):
        return self.col_offsets.size - 1

    def getNumberOfElements(self):
        return self.n_alignments

    @property
    def n_species(self):
        return self.codes.shape[1]

    def lengths(self):
        return np.diff(self.col_offsets)

    def getElementAt(self, i):
        return self.codes[self.col_offsets[i]:self.col_offsets[i + 1]]

    def subset(self, idx):
        idx = list(idx)
        if not idx:
            return PhyloSample(np.zeros((0, self.n_species), np.uint8), np.zeros(1, np.int64), species=self.species,
                               genes=[] if self.genes else None)
        return PhyloSample.from_alignments([self.getElementAt(i) for i in idx], species=self.species,
                                           genes=[self.genes[i] for i in idx] if self.genes else None)

    def slice(self, lo, hi):
        """Alignments lo..hi-1 as a sample of their own (one contiguous copy of the column block).  An empty range keeps
        the number of species."""
        lo, hi = int(lo), int(hi)
        c0, c1 = int(self.col_offsets[lo]), int(self.col_offsets[hi])
        out = PhyloSample(self.codes[c0:c1], self.col_offsets[lo:hi + 1] - c0, species=self.species,
                          genes=self.genes[lo:hi] if self.genes else None,
                          annotations=self.annotations[lo:hi] if self.annotations else None)
        out.learnedTopology = self.learnedTopology[lo:hi] if self.learnedTopology else None
        return out

    def shard(self, rank, world):
        """Contiguous block of alignments balanced by column count; every rank gets at least one
        alignment when there are at least `world` of them (sharding.shard_bounds)."""
        from .sharding import shard_bounds
        lo, hi = shard_bounds(self.col_offsets, world)[rank]
        return self.slice(lo, hi), (lo, hi)

    def device(self, ctx=None):
        """The packed, device-resident copy (created once per context)."""
        ctx = ctx or core.default_context()
        key = id(ctx)
        if key not in self._device:
            self._device[key] = core.Dataset(ctx, self.col_offsets, self.codes)
        return self._device[key]
