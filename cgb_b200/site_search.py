"""Thresholded binding-site search and the per-gene posterior loop.

Replaces the two per-gene Python loops of the reference with one tiled scan and
one batched posterior launch per genome:

  identify_sites                      cgb/genome.py:394-455   (HOT LOOP C)
  calculate_regulation_probabilities  cgb/genome.py:330-333, cgb/gene.py:129-146 (HOT LOOP B)

The reference scores each gene's upstream region separately (twice: the region
and its reverse complement); here every replicon is scanned once on the GPU and
the genome-wide hit list is joined with the gene regions on the host.  A window
belongs to a region iff it lies fully inside it, exactly the windows
``score_seq(region)`` enumerates; a site under two overlapping regions is reported
once per gene, as in the reference.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from .gene_regions import python_slice_bounds

# cgb/genome.py:22 — chromid is the replicon index in the PackedGenome, gene the
# row of the GeneTable
Site = namedtuple('Site', 'chromid start end strand score gene')


def identify_sites(model, genome, gene_tables, promoter_up_distance=300, promoter_dw_distance=50,
                   use_up_dist_site_scan=False, threshold=None, strict=True):
    """genome.py:394-448: sites with score >= threshold in the upstream region of
    every gene, both strands, sorted by score (stable, descending).

    ``gene_tables``: one GeneTable per replicon that has genes.  Returns the list
    of :class:`Site` in the reference's order.  ``strict``: a region more than one base shorter
    than the motif raises ValueError, as the reference does inside Biopython (e.g. a
    reverse-strand gene within ``promoter_dw_distance`` of the replicon start, whose region
    start goes negative and wraps); ``strict=False`` leaves such regions out."""
    if threshold is None:
        threshold = model.threshold()
    m = model.length
    seq, pos, strand, score = model.scan_hits(genome, threshold, revseq_order=True)
    parts = []
    for gt in gene_tables:
        k = gt.seq_index
        L = genome.length(k)
        if use_up_dist_site_scan:
            start, end = gt.upstream_noncoding_region_location(promoter_up_distance, promoter_dw_distance)
        else:
            start, end = gt.upstream_noncoding_region_location(None, promoter_dw_distance)
        s, e = python_slice_bounds(start, end, L)
        if strict and len(s) and int((e - s).min()) < m - 1:
            # score_seq on a region more than one base shorter than the motif: Biopython's
            # numpy.empty(len(seq) - m + 1) raises (a region of m - 1 bases has no window, no error)
            raise ValueError("negative dimensions are not allowed")
        sel = seq == k
        hp, hs, hv = pos[sel], strand[sel], score[sel]       # sorted by (pos, + strand first)
        # windows fully inside [s, e): s <= p <= e - m
        lo = np.searchsorted(hp, s, side="left")
        hi = np.maximum(np.searchsorted(hp, e - m, side="right"), lo)
        cnt = hi - lo
        total = int(cnt.sum())
        if total == 0:
            continue
        # one row per (gene, hit under its region), no per-gene loop: gene ids repeated by their
        # hit counts, hit indices as a ramp inside each gene's [lo, hi)
        gene = np.repeat(np.arange(len(cnt)), cnt)
        first = np.repeat(np.cumsum(cnt) - cnt, cnt)
        idx = np.repeat(lo, cnt) + (np.arange(total) - first)
        st = hs[idx]
        # the reference's order before its sort: genes in order, per gene the + strand hits by
        # position, then the - strand hits by position (genome.py:435-445)
        order = np.lexsort((idx, st == -1, gene))
        gene, idx, st = gene[order], idx[order], st[order]
        # reported coordinate: enumerate(scores, start=start) uses the UNclamped region start
        label = hp[idx] - s[gene] + start[gene]
        parts.append((np.full(total, k, dtype=np.int64), label, st, hv[idx], gene))
    if not parts:
        return []
    chrom, label, st, sc, gene = [np.concatenate(x) for x in zip(*parts)]
    # sites.sort(key=score, reverse=True) is stable: equal scores keep their order
    order = np.argsort(-sc.astype(np.float64), kind="stable")
    chrom, label, st, sc, gene = chrom[order], label[order], st[order], sc[order], gene[order]
    return [Site(c, a, a + m, t, v, g)
            for c, a, t, v, g in zip(chrom.tolist(), label.tolist(), st.tolist(), list(sc), gene.tolist())]


def calculate_regulation_probabilities(model, genome, gene_table, prior, alpha,
                                       promoter_up_distance=300, promoter_dw_distance=50):
    """genome.py:330-333 + gene.py:129-146: posterior probability of regulation
    for every gene of one replicon (device float64 tensor, gene order)."""
    start, end = gene_table.promoter_region_location(promoter_up_distance, promoter_dw_distance)
    return model.binding_probabilities(genome, start, end, prior, alpha, seq_index=gene_table.seq_index)


# ---------------------------------------------------------------------------
# identified_sites report
# ---------------------------------------------------------------------------
CSV_HEADER = ['score', 'site', 'chromid', 'start', 'end', 'strand', 'distance', 'category',
              'gene_strand', 'gene_start', 'gene_end', 'gene_locus_tag', 'protein_id', 'gene_product']

# Bio.Seq complement table (IUPAC ambiguity codes included, case kept)
_COMPLEMENT = str.maketrans("ACGTMRWSYKVHDBNacgtmrwsykvhdbn", "TGCAKYWSRMBDHVNtgcakywsrmbdhvn")


def relative_distance_to_start(gene_start, gene_end, gene_strand, site_start, site_end):
    """Gene.relative_distance_to_start (cgb/gene.py:290-316): distance of the site's distal end
    to the gene start; negative = upstream."""
    if gene_strand == 1:
        if site_start <= gene_start:
            return site_start - gene_start
        return site_end - gene_start
    if site_end >= gene_end:
        return gene_end - site_end
    return gene_end - site_start


def identified_sites_rows(sites, sequences, gene_tables, promoter_up_distance, accessions=None):
    """Rows of the reference's ``identified_sites/<genome>.csv`` (cgb/genome.py:457-494), one per
    Site, in the order given.  ``sequences[k]``: replicon k as str (for the site sequence,
    Chromid.subsequence incl. the reverse complement of minus-strand sites, chromid.py:92-97);
    ``gene_tables``: the GeneTables handed to identify_sites (optional ``locus_tag`` /
    ``protein_id`` / ``product`` lists on them fill the last columns)."""
    by_seq = {gt.seq_index: gt for gt in gene_tables}
    rows = []
    for site in sites:
        gt = by_seq[site.chromid]
        gi = site.gene
        g_start, g_end, g_strand = int(gt.start[gi]), int(gt.end[gi]), int(gt.strand[gi])
        dist = relative_distance_to_start(g_start, g_end, g_strand, site.start, site.end)
        # all downstream sites are "operator"; upstream ones beyond promoter_up are "intergenic"
        category = 'operator'
        if dist < 0:
            if abs(dist) > promoter_up_distance:
                category = 'intergenic'
        seq = sequences[site.chromid][site.start:site.end]
        if site.strand == -1:
            seq = seq.translate(_COMPLEMENT)[::-1]
        meta = lambda name: (getattr(gt, name)[gi] if getattr(gt, name, None) is not None else '')
        rows.append([site.score, seq, accessions[site.chromid] if accessions else site.chromid,
                     site.start, site.end, site.strand, dist, category, g_strand, g_start, g_end,
                     meta('locus_tag'), meta('protein_id'), meta('product')])
    return rows


def write_identified_sites_csv(filename, sites, sequences, gene_tables, promoter_up_distance, accessions=None):
    """Genome._output_identified_sites (cgb/genome.py:457-494)."""
    import csv
    with open(filename, 'w', newline='') as csvfile:
        w = csv.writer(csvfile)
        w.writerow(CSV_HEADER)
        for row in identified_sites_rows(sites, sequences, gene_tables, promoter_up_distance, accessions):
            w.writerow(row)
