"""Gene coordinate rules of the reference, vectorised over a gene table.

Restates (as int64 array arithmetic, no per-gene Python objects):
  Gene.upstream_gene                        cgb/gene.py:57-70
  Gene.upstream_noncoding_region_location   cgb/gene.py:72-104
  Gene.promoter_region                      cgb/gene.py:111-127
  Chromid.subsequence (forward strand)      cgb/chromid.py:92-97  (Python slice semantics)

The results are descriptor tables ``(start, end)`` that the GPU kernels consume;
the reference builds one Python ``str`` per gene instead.
"""
from __future__ import annotations

import numpy as np


class GeneTable:
    """Genes of one replicon in annotation order: start, end (0-based, end
    exclusive, as Biopython FeatureLocation) and strand (+1 / -1)."""

    def __init__(self, start, end, strand, chrom_len, seq_index=0, locus_tag=None, protein_id=None,
                 product=None, index=None, name=None, product_type=None):
        self.start = np.asarray(start, dtype=np.int64)
        self.end = np.asarray(end, dtype=np.int64)
        self.strand = np.asarray(strand, dtype=np.int64)
        self.chrom_len = int(chrom_len)
        self.seq_index = int(seq_index)
        # optional annotation columns of the identified_sites report (genome.py:470-493)
        self.locus_tag, self.protein_id, self.product = locus_tag, protein_id, product
        self.name, self.product_type = name, product_type
        if not (len(self.start) == len(self.end) == len(self.strand)):
            raise ValueError("start/end/strand must have equal lengths")
        # Gene._index (chromid.py:101-121): the running number of the 'gene' FEATURE, which
        # keeps counting over the compound-location genes the reference leaves out of the
        # list.  Gene.upstream_gene indexes the (shorter) list with it, so after a skipped
        # gene "the previous gene" is really a later one.  Default: list position.
        self.index = np.arange(len(self.start), dtype=np.int64) if index is None else np.asarray(index, dtype=np.int64)
        if len(self.index) != len(self.start):
            raise ValueError("index must have one entry per gene")

    def __len__(self):
        return len(self.start)

    # gene.py:57-70: genes[_index - 1] for forward genes with _index > 0, genes[_index + 1] for
    # reverse genes with _index < len(genes) - 1, else None
    def _upstream(self):
        n = len(self)
        fwd = self.strand == 1
        idx = np.where(fwd, self.index - 1, self.index + 1)
        has = np.where(fwd, self.index > 0, self.index < n - 1)
        if np.any(has & (idx >= n)):
            # the reference's list lookup fails the same way (forward gene after enough
            # skipped compound genes)
            raise IndexError("list index out of range (Gene.upstream_gene, cgb/gene.py:66)")
        idx = np.clip(idx, 0, max(n - 1, 0))
        return idx, has

    def upstream_gene(self):
        """Row of each gene's upstream gene (gene.py:57-70), -1 where there is none."""
        idx, has = self._upstream()
        return np.where(has, idx, -1)

    def upstream_noncoding_region_location(self, up=None, down=50):
        """gene.py:72-104.  ``if up:`` means 0 behaves like None."""
        fwd = self.strand == 1
        if up:
            f_start = np.maximum(0, self.start - up)
            r_end = np.minimum(self.chrom_len, self.end + up)
        else:
            uidx, has = self._upstream()          # only looked up on this branch (gene.py:85-88)
            f_start = np.where(has, np.minimum(self.end[uidx] - down, self.start), 0)
            r_end = np.where(has, np.maximum(self.start[uidx] + down, self.end), self.chrom_len)
        f_end = self.start + down
        r_start = self.end - down
        return np.where(fwd, f_start, r_start), np.where(fwd, f_end, r_end)

    def promoter_region_location(self, up=300, down=50):
        """gene.py:111-127 (bounds later handed to a Python slice)."""
        fwd = self.strand == 1
        f_start = np.maximum(0, self.start - up)
        f_end = self.start + down
        r_start = self.end - down
        r_end = np.minimum(self.chrom_len, self.end + up)
        return np.where(fwd, f_start, r_start), np.where(fwd, f_end, r_end)


def python_slice_bounds(start, end, length):
    """``sequence[start:end]`` -> effective [s, e) with 0 <= s <= e <= length
    (negative indices wrap once, then clamp; chromid.py:94)."""
    start = np.asarray(start, dtype=np.int64)
    end = np.asarray(end, dtype=np.int64)
    s = np.where(start < 0, np.maximum(start + length, 0), np.minimum(start, length))
    e = np.where(end < 0, np.maximum(end + length, 0), np.minimum(end, length))
    e = np.maximum(e, s)
    return s, e
