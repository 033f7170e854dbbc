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
    of every directon with more than one gene, over all replicons.  (The distances are integers:
    their sum is exact in any order, misc.mean's left-to-right sum / float(len) included.)"""
    total, count = 0, 0
    for gt in gene_tables:
        n = len(gt)
        if n < 2:
            continue
        st = np.asarray(gt.start, dtype=np.int64)
        en = np.asarray(gt.end, dtype=np.int64)
        sd = np.asarray(gt.strand)
        order = np.argsort(st, kind="stable")
        so = sd[order]
        run_start = np.concatenate([[0], np.nonzero(so[1:] != so[:-1])[0] + 1])
        run_end = np.append(run_start[1:], n)
        multi = run_end - run_start > 1
        fwd = so[run_start] == 1
        # first two genes of the walk: the run's first two, or (walked backwards) its last two
        a = np.where(fwd, run_start, run_end - 1)[multi]
        b = np.where(fwd, run_start + 1, run_end - 2)[multi]
        ga, gb = order[a], order[b]
        d = np.maximum(st[ga], st[gb]) - np.minimum(en[ga], en[gb])
        total += int(d.sum())
        count += int(len(d))
    return sigma * (total / float(count))


def predict_operons(gt, distance_th, probability_th, regulation_probability, start_id=1, accession=None):
    """Chromid.operon_prediction for one replicon.  ``regulation_probability``: per-gene posterior
    (array in GeneTable order); genes above ``probability_th`` start a new operon unless
    ``probability_th == 1.0`` (splitting off).  Returns ``(operons, n_splits)``.

    Array form of the reference's ``while directons_rest`` loop (chromid.py:216-232).  The loop walks
    every directon once per pass and peels one operon off its front; a pair of neighbouring genes is
    therefore examined exactly once, an operon is a maximal run of a directon without a break (the
    distance to the previous gene at or above the threshold, or -- the distance being below it -- a
    posterior above the threshold), and the operons come out ordered by (pass = index of the run in
    its directon, directon).  All of that is computed with array operations; ids, gene order, splits
    counted and every field are those of the loop (``predict_operons_loop``, compared in the tests)."""
    n = len(gt)
    if n == 0:
        return [], 0
    st = np.asarray(gt.start, dtype=np.int64)
    en = np.asarray(gt.end, dtype=np.int64)
    sd = np.asarray(gt.strand)
    if probability_th != 1.0:
        split = np.asarray(regulation_probability) > probability_th
    else:
        split = np.zeros(n, dtype=bool)
    # directons: runs of one strand along the (stable) start order; a run whose strand is not +1 is
    # walked from its last gene backwards
    order = np.argsort(st, kind="stable")
    so = sd[order]
    new_run = np.empty(n, dtype=bool)
    new_run[0] = True
    new_run[1:] = so[1:] != so[:-1]
    run_of = np.cumsum(new_run) - 1
    run_start = np.nonzero(new_run)[0]
    run_len = np.diff(np.append(run_start, n))
    pos = np.arange(n)
    k = pos - run_start[run_of]
    backwards = so[run_start][run_of] != 1
    walk = order[np.where(backwards, run_start[run_of] + run_len[run_of] - 1 - k, pos)]   # genes in loop order
    # breaks between neighbours of a walk (Gene.distance, gene.py:283-288)
    a, b = walk[:-1], walk[1:]
    far = (np.maximum(st[a], st[b]) - np.minimum(en[a], en[b])) >= distance_th
    brk = np.ones(n, dtype=bool)
    brk[1:] = new_run[1:] | far | split[b]
    splits = int(np.count_nonzero(~new_run[1:] & ~far & split[b]))
    seg_of = np.cumsum(brk) - 1                         # operon (segment) of every walked gene
    seg_start = np.nonzero(brk)[0]
    n_seg = len(seg_start)
    seg_run = run_of[seg_start]
    first_seg_of_run = np.searchsorted(seg_run, np.arange(len(run_start)))
    seg_pass = np.arange(n_seg) - first_seg_of_run[seg_run]
    seg_order = np.lexsort((seg_run, seg_pass))         # the loop's order: pass, then directon
    rank_of_seg = np.empty(n_seg, dtype=np.int64)
    rank_of_seg[seg_order] = np.arange(n_seg)
    # genes of every operon sorted by start, ties in walk order (sorted() is stable, operon.py:26)
    g_order = np.lexsort((pos, st[walk], rank_of_seg[seg_of]))
    genes_sorted = walk[g_order]
    counts = np.diff(np.append(seg_start, n))[seg_order]
    bounds = np.concatenate([[0], np.cumsum(counts)])
    o_start = np.minimum.reduceat(st[genes_sorted], bounds[:-1])
    o_end = np.maximum.reduceat(en[genes_sorted], bounds[:-1])
    o_strand = sd[genes_sorted[bounds[:-1]]]
    o_first = np.where(o_strand == 1, genes_sorted[bounds[:-1]], genes_sorted[bounds[1:] - 1])   # operon.py:68-73
    prob = np.asarray(regulation_probability, dtype=np.float64)[o_first].tolist() \
        if regulation_probability is not None else [float("nan")] * n_seg
    chrom = accession if accession is not None else gt.seq_index
    gl, bl = genes_sorted.tolist(), bounds.tolist()
    seq_index = gt.seq_index
    out = [Operon(start_id + i, seq_index, chrom, gl[bl[i]:bl[i + 1]], s0, e0, sd0, f0, p0)
           for i, (s0, e0, sd0, f0, p0) in enumerate(zip(o_start.tolist(), o_end.tolist(), o_strand.tolist(),
                                                         o_first.tolist(), prob))]
    return out, splits


def predict_operons_loop(gt, distance_th, probability_th, regulation_probability, start_id=1, accession=None):
    """The reference's loop shape over plain lists (kept as the differential check of
    :func:`predict_operons`).  Chromid.operon_prediction for one replicon.  ``regulation_probability``: per-gene posterior
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
