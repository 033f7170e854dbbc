"""The C feature-table index (cgbk_genbank_index, host code of libcgbk) against the pure-Python
walk of the same table (genbank._parse_features): same Feature tuples -- types, locations,
qualifier dictionaries -- on hand-written tricky tables and on randomised ones, and the same
gene table whichever path builds it (array fast path, per-gene path, Feature path).
Reference: SeqIO.read(..., 'gb') + Chromid.genes, cgb/chromid.py:32-35,99-122."""
import random

import numpy as np
import pytest

from cgb_b200 import genbank
from cgb_b200.genbank import FeatureIndex, GenBankRecord, _parse_features

Q = " " * 21
TRICKY = "\n".join([
    "     source          1..5000",
    Q + '/organism="synthetic',
    Q + 'construct"',
    Q + '/mol_type="genomic DNA"',
    "     gene            complement(<10..>200)",
    Q + '/locus_tag="T_0001"',
    Q + "/pseudo",
    Q + '/note="a value with ""doubled"" quotes and a slash line',
    Q + '/that is still inside the quotes" ',
    "     CDS             complement(join(10..50,",
    Q + "60..200))",
    Q + '/locus_tag="T_0001"',
    Q + "/codon_start=1",
    Q + '/product="hypothetical',
    Q + 'protein"',
    Q + '/translation="MKKLLPIAVAGLLSACSHEELAAKLNQYQVRRDAIRERMA',
    Q + 'QLEDEIRRLKEALR"',
    "     gene            J01234.1:300..400",
    Q + '/locus_tag="T_0002"',
    Q + '/gene="abcD"',
    "     misc_feature    order(500..510,AB12.3:1..4,520..530)",
    Q + '/note="x"',
    "     gene            600^601",
    Q + '/locus_tag="T_0003"',
    "     tRNA            700",
    Q + '/locus_tag="T_0003"',
    Q + '/product="tRNA-Ala"',
    "     gene            800.900",
    Q + '/locus_tag="T_0004"',
    Q + '/locus_tag="T_0004b"',
    "     repeat_region   nonsense..location",
    Q + "/rpt_type=tandem",
    "     gene            1000..1200",
    Q + '/locus_tag="T_0005"',
    "     CDS             1000..1200",
    Q + '/locus_tag="T_0005"',
    Q + '/protein_id="WP_000000005.1"',
    Q + '/product=""',
]) + "\n"


def _features_from_index(block_text):
    ix = FeatureIndex(block_text.encode("ascii"))
    return [ix.feature(k) for k in range(len(ix))]


def test_tricky_table_matches_the_python_walk():
    got, exp = _features_from_index(TRICKY), _parse_features(TRICKY)
    assert len(got) == len(exp) == 11
    for g, e in zip(got, exp):
        assert g == e, (g, e)
    by_type = [f.type for f in got]
    assert by_type[1] == "gene" and got[1].strand == -1 and (got[1].start, got[1].end) == (9, 200)
    assert got[1].qualifiers["note"] == ['a value with "doubled" quotes and a slash line /that is still inside the quotes']
    assert got[1].qualifiers["pseudo"] == [""]
    assert got[2].compound and (got[2].start, got[2].end, got[2].strand) == (9, 200, -1)
    assert got[2].qualifiers["translation"] == ["MKKLLPIAVAGLLSACSHEELAAKLNQYQVRRDAIRERMAQLEDEIRRLKEALR"]
    assert got[2].qualifiers["product"] == ["hypothetical protein"]
    assert (got[3].start, got[3].end, got[3].compound) == (299, 400, False)      # accession prefix
    assert got[4].compound and (got[4].start, got[4].end) == (0, 530)             # hull of all numbers, accession names stripped
    assert (got[5].start, got[5].end) == (600, 600) and (got[6].start, got[6].end) == (699, 700)
    assert (got[7].start, got[7].end) == (799, 900)
    assert got[8].compound and (got[8].start, got[8].end) == (0, 0)               # unparsable location


def _random_table(rng, n_feat):
    lines = []
    pos = 1
    k = 0
    for _ in range(n_feat):
        ftype = rng.choice(["gene", "gene", "CDS", "tRNA", "misc_feature", "regulatory"])
        a = pos + rng.randint(0, 50)
        b = a + rng.randint(1, 900)
        pos = b
        form = rng.random()
        if form < 0.6:
            loc = "%s%d..%s%d" % (rng.choice(["", "", "<"]), a, rng.choice(["", "", ">"]), b)
        elif form < 0.75:
            mid = (a + b) // 2
            loc = "%s(%d..%d,%d..%d)" % (rng.choice(["join", "order"]), a, mid, mid + 1, b)
        elif form < 0.8:
            loc = "%d" % a
        elif form < 0.85:
            loc = "%d^%d" % (a, a + 1)
        else:
            loc = "ACC%d.1:%d..%d" % (rng.randint(1, 99), a, b)
        if rng.random() < 0.4:
            loc = "complement(%s)" % loc
        if len(loc) > 30 and "," in loc and rng.random() < 0.5:       # a location broken over two lines
            cut = loc.index(",") + 1
            lines.append("     %-15s %s" % (ftype, loc[:cut]))
            lines.append(Q + loc[cut:])
        else:
            lines.append("     %-15s %s" % (ftype, loc))
        if ftype == "gene" or rng.random() < 0.7:
            if ftype == "gene":
                k += 1
            lines.append(Q + '/locus_tag="R_%04d"' % k)
            if rng.random() < 0.05:
                lines.append(Q + '/locus_tag="R_%04d_alt"' % k)
        if rng.random() < 0.3:
            lines.append(Q + '/gene="g%d"' % rng.randint(1, 999))
        if ftype == "CDS" and rng.random() < 0.8:
            lines.append(Q + '/protein_id="WP_%09d.1"' % rng.randint(1, 10 ** 9 - 1))
        if rng.random() < 0.6:
            words = " ".join(rng.choice(["alpha", "beta/gamma", 'say ""hi""', "kinase", "/slashed", "x" * 30])
                             for _ in range(rng.randint(1, 14)))
            text = '/product="%s"' % words
            while len(text) > 58:                                        # wrapped value, '/' may open a line
                cut = text.rfind(" ", 0, 59)
                if cut <= 0:
                    cut = 58
                lines.append(Q + text[:cut])
                text = text[cut:].lstrip(" ")
            lines.append(Q + text)
        if rng.random() < 0.2:
            lines.append(Q + "/pseudo")
        if rng.random() < 0.2:
            lines.append(Q + "/codon_start=%d" % rng.randint(1, 3))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("seed", range(12))
def test_random_tables_match_the_python_walk(seed):
    rng = random.Random(seed)
    block = _random_table(rng, rng.randint(1, 120))
    got, exp = _features_from_index(block), _parse_features(block)
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert g == e, (g, e)
    # gene table: index paths (array fast path or per-gene) == Feature path
    seq = b"A" * 200000

    def table_of(rec):
        try:
            return rec.gene_table()
        except (KeyError, AssertionError) as exc:
            return type(exc).__name__
    a = table_of(GenBankRecord("X.1", "d", seq, FeatureIndex(block.encode("ascii"))))
    b = table_of(GenBankRecord("X.1", "d", seq, exp))
    if isinstance(a, str) or isinstance(b, str):
        assert a == b
        return
    for col in ("start", "end", "strand", "index"):
        assert np.array_equal(getattr(a, col), getattr(b, col)), col
    for col in ("locus_tag", "name", "product", "protein_id", "product_type"):
        assert list(getattr(a, col)) == list(getattr(b, col)), col


def test_empty_and_headerless_tables():
    assert _features_from_index("") == [] and _parse_features("") == []
    junk = "no feature here\n   nor here\n"
    assert _features_from_index(junk) == _parse_features(junk) == []


def _plain_records(text):
    """An independent reading of a multi-record flat file: records cut at the '//' lines, the ORIGIN
    block cleaned with str methods, the feature table bounded by FEATURES / ORIGIN / section words."""
    out = []
    for chunk in text.replace("\r\n", "\n").split("\n//"):
        if "LOCUS" not in chunk:
            continue
        chunk = chunk[chunk.index("LOCUS"):]
        seq = ""
        if "\nORIGIN" in chunk:
            head, origin = chunk.split("\nORIGIN", 1)
            origin = origin.split("\n", 1)[1] if "\n" in origin else ""
            seq = "".join(ch for ch in origin if ch not in " 0123456789\t\r\n").upper()
        else:
            head = chunk
        n_gene = 0
        if "\nFEATURES" in head:
            table = head.split("\nFEATURES", 1)[1]
            for word in ("\nCONTIG", "\nBASE COUNT", "\nWGS", "\nCOMMENT"):
                table = table.split(word)[0]
            n_gene = sum(1 for line in table.split("\n") if line.startswith("     gene "))
        out.append((seq, n_gene))
    return out


def test_origin_decoder_and_record_bounds():
    """cgbk_genbank_origin + the record cutting of genbank.parse against an independent reading:
    several records in one file, one without ORIGIN (CONTIG only), one without FEATURES, ragged and
    non-standard ORIGIN lines, lower / upper case, IUPAC codes, CRLF line ends."""
    import random

    from genbank_writer import genbank_text
    rng = random.Random(11)

    def rep(n, k):
        seq = "".join(rng.choice("ACGTNRYacgtn") for _ in range(n))
        feats = [{"type": "source", "start": 0, "end": n, "strand": 1, "compound": False, "qualifiers": {}}]
        for j in range(k):
            a = 10 + 90 * j
            feats.append({"type": "gene", "start": a, "end": a + 60, "strand": 1 if j % 2 else -1, "compound": False,
                          "qualifiers": {"locus_tag": ["T%d_%03d" % (n, j)]}})
            feats.append({"type": "CDS", "start": a, "end": a + 60, "strand": 1 if j % 2 else -1, "compound": False,
                          "qualifiers": {"locus_tag": ["T%d_%03d" % (n, j)], "product": ["protein // not a terminator"]}})
        return {"accession": "R%d.1" % n, "description": "record of %d" % n, "sequence": seq, "features": feats}

    a, b, c, d = (genbank_text(rep(n, k)) for n, k in ((1234, 6), (600, 3), (61, 0), (7, 0)))
    # b: ORIGIN replaced by a CONTIG section (no sequence); c: no feature table at all;
    # d: a hand-written ORIGIN with blocks of 5 and trailing blanks
    b = b[:b.index("ORIGIN")] + "CONTIG      join(XX000001.1:1..600)\n//\n"
    c = c[:c.index("FEATURES")] + c[c.index("ORIGIN"):]
    d = d[:d.index("ORIGIN")] + "BASE COUNT      2 a 2 c 2 g 1 t\nORIGIN      \n        1 acgtn ry  \n\n//\n"
    for text in (a + b + c + d, c + a, d + b + a, (a + b + c + d).replace("\n", "\r\n"), a[:-3], a.upper()):
        exp = _plain_records(text)
        for src in (text, text.encode("ascii")):
            got = genbank.parse(src)
            assert [(r.sequence.decode("ascii"), sum(1 for f in r.features if f.type == "gene")) for r in got] == exp
    recs = genbank.parse(a + b + c + d)
    assert [r.id for r in recs] == ["R1234.1", "R600.1", "R61.1", "R7.1"] and recs[3].sequence == b"ACGTNRY"
    assert len(recs[1].sequence) == 0 and len(recs[2].features) == 0
    # the decoder alone: capacity errors and the end offset
    import ctypes as C
    from cgb_b200 import _cabi
    lib = _cabi.lib()
    blob = b"        1 acgtacgtac gtacgtacgt acgtacgtac gtacgtacgt acgtacgtac gtacgtacgt\n       61 ac\n//\nLOCUS"
    out = np.zeros(80, dtype=np.uint8)
    nb, end = C.c_int64(), C.c_int64()
    assert lib.cgbk_genbank_origin(blob, len(blob), 0, out.ctypes.data, 80, C.byref(nb), C.byref(end)) == 0
    assert nb.value == 62 and out[:62].tobytes() == b"ACGT" * 15 + b"AC" and blob[end.value:end.value + 2] == b"//"
    assert lib.cgbk_genbank_origin(blob, len(blob), 0, out.ctypes.data, 61, C.byref(nb), C.byref(end)) != 0
