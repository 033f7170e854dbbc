"""Property tests (hypothesis) of the small host-side rules the kernels' inputs depend on."""
import numpy as np
from hypothesis import given, settings, strategies as st

from cgb_b200 import GeneTable, genbank, operons, python_slice_bounds


@settings(max_examples=300, deadline=None)
@given(st.integers(0, 60), st.integers(-90, 90), st.integers(-90, 90))
def test_python_slice_bounds_is_python_slicing(length, start, end):
    """Chromid.subsequence is ``sequence[start:end]`` (cgb/chromid.py:94): the effective bounds
    must select exactly what Python selects, for negative, overshooting and crossed indices."""
    s, e = python_slice_bounds(np.array([start]), np.array([end]), length)
    assert list(range(length))[start:end] == list(range(int(s[0]), int(e[0])))


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 10 ** 7), st.integers(0, 5000), st.booleans(), st.booleans(), st.booleans())
def test_location_round_trip(a, span, complement, fuzzy_lo, fuzzy_hi):
    b = a + span
    text = "%s%d..%s%d" % ("<" if fuzzy_lo else "", a, ">" if fuzzy_hi else "", b)
    if complement:
        text = "complement(%s)" % text
    assert genbank.parse_location(text) == (a - 1, b, -1 if complement else 1, False)
    joined = "join(%d..%d,%d..%d)" % (a, b, b + 10, b + 20)
    assert genbank.parse_location("complement(%s)" % joined if complement else joined)[2:] == \
        (-1 if complement else 1, True)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 5000), st.integers(1, 400), st.sampled_from([1, -1])), min_size=1,
                max_size=40), st.floats(0.0, 1.0), st.floats(0.2, 3.0), st.integers(0, 2 ** 31))
def test_operons_partition_the_genes(genes, th, sigma, seed):
    """Whatever the gene layout and the posteriors: every gene lands in exactly one operon, an
    operon is one strand, consecutive in start order, and its first gene is the upstream-most."""
    start = np.array([g[0] for g in genes]); end = start + np.array([g[1] for g in genes])
    strand = np.array([g[2] for g in genes])
    gt = GeneTable(start, end, strand, int(end.max()) + 10)
    gt.locus_tag = ["g%d" % i for i in range(len(gt))]; gt.product = [""] * len(gt)
    probs = np.random.default_rng(seed).random(len(gt))
    oprs, _ = operons.predict_operons(gt, 50.0 * sigma, th, probs)
    seen = sorted(g for o in oprs for g in o.genes)
    assert seen == list(range(len(gt)))
    order = {int(g): k for k, g in enumerate(np.argsort(start, kind="stable"))}
    for o in oprs:
        assert len({int(strand[g]) for g in o.genes}) == 1
        ranks = sorted(order[g] for g in o.genes)
        assert ranks == list(range(ranks[0], ranks[0] + len(ranks)))       # a run of neighbours
        assert o.first_gene == (o.genes[0] if o.strand == 1 else o.genes[-1])
        assert o.start == min(start[g] for g in o.genes) and o.end == max(end[g] for g in o.genes)
    assert [o.operon_id for o in oprs] == list(range(1, len(oprs) + 1))


def test_array_operon_prediction_equals_the_reference_loop_shape():
    """operons.predict_operons (array operations) against predict_operons_loop (the reference's
    ``while directons_rest`` loop over lists, chromid.py:216-232) on random gene tables: equal starts,
    strand 0, overlapping genes, every threshold regime.  Operon ids, gene order inside an operon,
    start / end / strand / first gene / probability and the number of splits must be identical."""
    import numpy as np

    from cgb_b200 import GeneTable
    from cgb_b200 import operons as op
    rng = np.random.default_rng(123)
    for trial in range(150):
        n = int(rng.integers(1, 70))
        L = 30000
        start = np.sort(rng.integers(0, L - 1000, size=n))
        if n > 4 and rng.random() < 0.4:
            start[1] = start[0]
            start[4] = start[3]
        rng.shuffle(start)
        end = start + rng.integers(30, 900, size=n)
        strand = rng.choice([1, -1], size=n)
        if rng.random() < 0.15:
            strand[int(rng.integers(0, n))] = 0
        gt = GeneTable(start.tolist(), end.tolist(), strand.tolist(), L, seq_index=int(rng.integers(0, 3)))
        prob = rng.random(n)
        for pth in (1.0, 0.5, 0.05):
            for dth in (50.0, 180.25, -20.0, 1e9):
                a, sa = op.predict_operons(gt, dth, pth, prob, start_id=5, accession="ACC")
                b, sb = op.predict_operons_loop(gt, dth, pth, prob, start_id=5, accession="ACC")
                assert a == b and sa == sb
                assert sorted(g for o in a for g in o.genes) == list(range(n))
        a, _ = op.predict_operons(gt, 100.0, 1.0, None)
        b, _ = op.predict_operons_loop(gt, 100.0, 1.0, None)
        assert [x[:8] for x in a] == [y[:8] for y in b] and all(np.isnan(x[8]) for x in a)
    assert op.predict_operons(GeneTable([], [], [], 10), 1.0, 1.0, None) == ([], 0)
