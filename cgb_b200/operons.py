"""Distance-based operon prediction with posterior-driven splitting.

Restates, over GeneTable arrays (no per-gene objects):
  Chromid.directons                      cgb/chromid.py:140-163
  Genome.intergenic_distance_threshold   cgb/genome.py:74-85   (+ misc.mean, cgb/misc.py:9-11)
  Chromid.operon_prediction              cgb/chromid.py:173-243
  Genome.operon_prediction / .operons    cgb/genome.py:87-116
  Operon                                 cgb/operon.py:22-86
  Genome.operons_to_csv                  cgb/genome.py:130-152
  Genome.infer_regulons / _output_posterior_probabilities   cgb/genome.py:335-378

The posteriors that drive the splitting come from the GPU path
(``calculate_regulation_probabilities``); everything here is host-side bookkeeping over a few
thousand genes and keeps the reference's order of operons (its ids depend on it).
"""
from __future__ import annotations

import csv
from collections import namedtuple

import numpy as np

# seq_index: replicon (GeneTable.seq_index); chromid: its accession for the reports; genes: rows
# of the replicon's GeneTable, sorted by start (operon.py:26)
Operon = namedtuple("Operon", "operon_id seq_index chromid genes start end strand first_gene "
                              "regulation_probability")


def gene_distance(gt, a, b):
    """Gene.distance (cgb/gene.py:283-288): intergenic distance, negative when the genes overlap."""
    return int(max(gt.start[a], gt.start[b]) - min(gt.end[a], gt.end[b]))


def directons(gt):
    """Chromid.directons: runs of consecutive same-strand genes in start order (stable sort),
    reverse-strand runs listed from their last gene backwards.  Lists of GeneTable rows."""
    if len(gt) == 0:
        return []
    order = np.argsort(gt.start, kind="stable")
    # run boundaries where the strand changes along the start order (array operations; the runs
    # themselves are short Python lists of GeneTable rows)
    strand = np.asarray(gt.strand)[order]
    cuts = (np.nonzero(strand[1:] != strand[:-1])[0] + 1).tolist()
    rows = order.tolist()
    first = strand[[0] + cuts].tolist()
    runs = [rows[a:b] for a, b in zip([0] + cuts, cuts + [len(rows)])]
    return [r if s == 1 else r[::-1] for r, s in zip(runs, first)]


def intergenic_distance_threshold(gene_tables, sigma=1.0):
    """Genome.intergenic_distance_threshold: sigma * mean distance between the first two genes
    of every directon with more than one gene, over all replicons."""
    dists = []
    for gt in gene_tables:
        st, en = np.asarray(gt.start).tolist(), np.asarray(gt.end).tolist()
        dists.extend(max(st[d[0]], st[d[1]]) - min(en[d[0]], en[d[1]]) for d in directons(gt) if len(d) > 1)
    total = 0
    for x in dists:                      # misc.mean: left-to-right sum / float(len)
        total = total + x
    return sigma * (total / float(len(dists)))


def predict_operons(gt, distance_th, probability_th, regulation_probability, start_id=1, accession=None):
    """Chromid.operon_prediction for one replicon.  ``regulation_probability``: per-gene posterior
    (array in GeneTable order); genes above ``probability_th`` start a new operon unless
    ``probability_th == 1.0`` (splitting off).  Returns ``(operons, n_splits)``."""
    if probability_th != 1.0:
        split = np.asarray(regulation_probability) > probability_th
    else:
        split = np.zeros(len(gt), dtype=bool)
    groups, splits = [], 0
    rest = directons(gt)
    # plain lists: the loops below index single elements, which numpy arrays are slow at
    split = split.tolist()
    st, en, sd = np.asarray(gt.start).tolist(), np.asarray(gt.end).tolist(), np.asarray(gt.strand).tolist()
    while rest:
        processing, rest = rest, []
        for d in processing:
            operon = [d[0]]
            i = 1
            while i < len(d):
                a, b = d[i - 1], d[i]
                if max(st[a], st[b]) - min(en[a], en[b]) >= distance_th:       # Gene.distance, gene.py:283-288
                    break
                if split[d[i]]:
                    splits += 1
                    break
                operon.append(d[i])
                i += 1
            groups.append(operon)
            if i < len(d):
                rest.append(d[i:])
    out = []
    prob = np.asarray(regulation_probability, dtype=np.float64).tolist() if regulation_probability is not None else None
    chrom = accession if accession is not None else gt.seq_index
    for oid, genes in enumerate(groups, start=start_id):
        if len(genes) > 1:
            genes = sorted(genes, key=st.__getitem__)                  # operon.py:26 (stable)
        strand = sd[genes[0]]
        first = genes[0] if strand == 1 else genes[-1]                 # operon.py:68-73
        out.append(Operon(oid, gt.seq_index, chrom, genes, min(st[g] for g in genes), max(en[g] for g in genes),
                          strand, first, prob[first] if prob is not None else float("nan")))
    return out, splits


def predict_operons_genome(gene_tables, probability_th, sigma, regulation_probabilities, accessions=None):
    """Genome.operon_prediction: replicon after replicon, ids running on (genome.py:101-112)."""
    distance_th = intergenic_distance_threshold(gene_tables, sigma)
    operons, start_id = [], 1
    for k, gt in enumerate(gene_tables):
        oprs, _ = predict_operons(gt, distance_th, probability_th, regulation_probabilities[k], start_id,
                                  accessions[k] if accessions else None)
        if oprs:
            start_id = oprs[-1].operon_id + 1
        operons.extend(oprs)
    return operons


def _tags_products(opr, gt):
    """'locus_tags' and 'products' cells of an operon row.  Both reports (operons, regulons) ask for
    every operon of the genome: the per-gene strings are built once per gene table and kept on it."""
    cells = gt.__dict__.get("_report_cells")
    if cells is None:
        lt, pr = gt.locus_tag, gt.product
        cells = gt.__dict__["_report_cells"] = (lt, [t + ' (%s)' % p for t, p in zip(lt, pr)])
    lt, lp = cells
    genes = opr.genes
    if len(genes) == 1:
        g = genes[0]
        return lt[g], lp[g]
    return ', '.join([lt[g] for g in genes]), ', '.join([lp[g] for g in genes])


def operons_rows(operons, gene_tables):
    """Rows of Genome.operons_to_csv (header first)."""
    by = {gt.seq_index: gt for gt in gene_tables}
    rows = [['operon_id', 'chromid', 'start', 'end', 'strand', 'locus_tags', 'products']]
    for opr in operons:
        tags, prods = _tags_products(opr, by[opr.seq_index])
        rows.append([opr.operon_id, opr.chromid, opr.start, opr.end, opr.strand, tags, prods])
    return rows


def infer_regulons(operons, threshold=0.5):
    """Genome.infer_regulons: ``(results, regulons)`` -- every operon with its posterior sorted
    by it (descending, stable), and the operons at or above ``threshold``."""
    results = [(opr, opr.regulation_probability) for opr in operons]
    results.sort(key=lambda x: x[1], reverse=True)
    return results, [opr for opr, p in results if p >= threshold]


def regulon_rows(results, gene_tables):
    """Rows of Genome._output_posterior_probabilities (header first)."""
    by = {gt.seq_index: gt for gt in gene_tables}
    rows = [['probability', 'operon_start', 'operon_end', 'operon_strand', 'genes', 'products']]
    for opr, prob in results:
        tags, prods = _tags_products(opr, by[opr.seq_index])
        rows.append(['%.3f' % prob, opr.start, opr.end, opr.strand, tags, prods])
    return rows


def write_rows(filename, rows):
    with open(filename, 'w', newline='') as f:
        csv.writer(f).writerows(rows)
