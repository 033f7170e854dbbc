"""GenBank flat file -> replicon sequence + gene table.

The reference gets a Biopython ``SeqRecord`` (``SeqIO.read(..., 'gb')``, cgb/chromid.py:32-35) and
turns its feature table into ``Gene`` objects (``Chromid.genes``, cgb/chromid.py:99-122).  Here one
pass over the text yields what the scan kernels and the reports need and nothing else:

  * the record id (VERSION, as Biopython's ``record.id``), the description and the sequence as
    upper-case ASCII bytes (``str(record.seq)``), ready for ``PackedGenome.from_sequences``;
  * the feature table as plain tuples, and from it a :class:`~cgb_b200.gene_regions.GeneTable`
    with the reference's selection rules:
      - only ``gene`` features, each paired with the feature that follows it; the product feature
        is that follower iff both carry the same ``/locus_tag`` list (chromid.py:104-106);
      - the very last feature is never looked at (``zip(features, features[1:])``, chromid.py:103);
      - genes whose location is not one contiguous span (``join``/``order``) are left out, but
        still advance the running index (chromid.py:107-117) -- ``GeneTable.index`` keeps that
        number because ``Gene.upstream_gene`` indexes the shortened list with it;
      - start/end 0-based half-open, ``<``/``>`` fuzzy ends dropped (Biopython ``.position``).

Biopython is not a dependency (it is not installed here): the parser follows the published
INSDC feature-table layout (feature key in column 6, location/qualifiers from column 22).
"""
from __future__ import annotations

import re
from collections import namedtuple


from .gene_regions import GeneTable

# location: (start, end, strand) for one contiguous span; ``compound`` marks join/order/unparsable
Feature = namedtuple("Feature", "type start end strand compound qualifiers")

_SPAN = re.compile(r"^(?:[A-Za-z0-9_.]+:)?[<>]?(\d+)(?:\.\.[<>]?(\d+)|\.[<>]?(\d+)|\^(\d+))?$")
_UPPER = bytes.maketrans(b"abcdefghijklmnopqrstuvwxyz", b"ABCDEFGHIJKLMNOPQRSTUVWXYZ")


class GenBankRecord(object):
    """One LOCUS ... // entry."""

    def __init__(self, accession, description, sequence, features):
        self.id = accession
        self.description = description
        self.sequence = sequence            # bytes, upper case
        self.features = features            # [Feature]

    @property
    def length(self):
        return len(self.sequence)

    def gene_table(self, seq_index=0):
        """``Chromid.genes`` (cgb/chromid.py:99-122) as a GeneTable with the annotation columns
        the reports use (locus_tag, gene name, product type, product, protein_id)."""
        start, end, strand, index = [], [], [], []
        locus, name, ptype, product, protein = [], [], [], [], []
        k = 0
        for f, nxt in zip(self.features, self.features[1:]):
            if f.type != "gene":
                continue
            tags = f.qualifiers["locus_tag"]                      # KeyError as in the reference
            product_f = nxt if tags == nxt.qualifiers.get("locus_tag") else None
            if f.compound:
                k += 1
                continue
            assert len(tags) == 1                                 # Gene.locus_tag, gene.py:178
            start.append(f.start); end.append(f.end); strand.append(f.strand); index.append(k)
            locus.append(tags[0])
            name.append(f.qualifiers["gene"][0] if "gene" in f.qualifiers else tags[0])   # gene.py:166-173
            pt = product_f.type if product_f is not None else ""
            ptype.append(pt)
            product.append(product_f.qualifiers["product"][0]
                           if product_f is not None and "product" in product_f.qualifiers else "")
            # Genome._output_identified_sites asks for the protein id of CDS genes only; a CDS
            # without /protein_id yields '' (gene.py:202-211)
            protein.append(product_f.qualifiers["protein_id"][0]
                           if pt == "CDS" and "protein_id" in product_f.qualifiers else "")
            k += 1
        return GeneTable(start, end, strand, self.length, seq_index=seq_index, locus_tag=locus,
                         protein_id=protein, product=product, index=index, name=name, product_type=ptype)


def parse_location(text):
    """INSDC location string -> (start, end, strand, compound)."""
    s = "".join(text.split())
    strand = 1
    while s.startswith("complement(") and s.endswith(")"):
        s = s[len("complement("):-1]
        strand = -strand
    if s.startswith(("join(", "order(", "bond(")):
        # CompoundLocation: report the hull; the gene rules only need to know it is not simple
        nums = [int(x) for x in re.findall(r"\d+", re.sub(r"[A-Za-z0-9_.]+:", "", s))]
        return (min(nums) - 1 if nums else 0, max(nums) if nums else 0, strand, True)
    mt = _SPAN.match(s)
    if not mt:
        return (0, 0, strand, True)                  # Biopython: location None -> not a FeatureLocation
    a = int(mt.group(1))
    if mt.group(2) is not None:                      # a..b
        return (a - 1, int(mt.group(2)), strand, False)
    if mt.group(3) is not None:                      # a.b (one base somewhere in the range)
        return (a - 1, int(mt.group(3)), strand, False)
    if mt.group(4) is not None:                      # a^b (between two bases)
        return (a, a, strand, False)
    return (a - 1, a, strand, False)                 # single base


def _quote_closed(value):
    """A quoted qualifier value is complete once it ends with a quote and holds an even number
    of them (a literal quote inside the text is written as two)."""
    return len(value) >= 2 and value.endswith('"') and value.count('"') % 2 == 0


_FEATURE_START = re.compile(r"^ {5}(?=\S)", re.M)
_QUALIFIER_START = re.compile(r"\n {21}/")
_CONTINUATION = re.compile(r"\n\s*")


def _parse_features(block):
    """Feature table text (without its header line) -> [Feature].  The table is cut at feature
    keys (column 6) and at qualifier starts (column 22, '/') by regular expressions; a '/' that
    opens a continuation line INSIDE a quoted value is glued back."""
    feats = []
    for chunk in _FEATURE_START.split(block)[1:]:
        chunk = chunk.rstrip()
        parts = _QUALIFIER_START.split(chunk)
        head = parts[0]
        sp = head.find(" ")
        if sp < 0:
            ftype, loc = head.strip(), ""
        else:
            ftype, loc = head[:sp], head[sp:]
        start, end, strand, compound = parse_location(loc)
        quals = {}
        i, n = 1, len(parts)
        while i < n:
            q = parts[i]
            i += 1
            eq = q.find("=")
            if eq < 0:
                quals.setdefault(q.strip(), []).append("")          # flag qualifier (/pseudo)
                continue
            key, value = q[:eq], q[eq + 1:]
            if value.startswith('"'):
                while not _quote_closed(value.rstrip()) and i < n:  # a '/' line inside the quotes
                    value += "\n" + " " * 21 + "/" + parts[i]
                    i += 1
            if "\n" in value:
                value = _CONTINUATION.sub("" if key == "translation" else " ", value)
            value = value.rstrip()
            if value.startswith('"'):
                value = value[1:-1] if value.endswith('"') and len(value) >= 2 else value[1:]
                if '""' in value:
                    value = value.replace('""', '"')
            if key == "translation" and " " in value:
                value = value.replace(" ", "")
            quals.setdefault(key, []).append(value)
        feats.append(Feature(ftype, start, end, strand, compound, quals))
    return feats


def parse(text):
    """All records of a GenBank flat file (``str`` or ``bytes``).  The three big blocks of a
    record (header, feature table, ORIGIN) are cut out with ``find`` and the sequence is
    cleaned with one ``bytes.translate``: only the feature table is walked line by line."""
    if isinstance(text, str):
        text = text.encode("ascii", "replace")
    records = []
    pos, n = 0, len(text)
    while True:
        start = 0 if text.startswith(b"LOCUS", pos) and pos == 0 else text.find(b"\nLOCUS", pos)
        if start < 0:
            break
        if text[start:start + 1] == b"\n":
            start += 1
        end = text.find(b"\n//", start)
        if end < 0:
            end = n
        rec = text[start:end]
        pos = end + 3
        f0 = rec.find(b"\nFEATURES")
        o0 = rec.find(b"\nORIGIN")
        header = rec[:f0 if f0 >= 0 else (o0 if o0 >= 0 else len(rec))].decode("ascii", "replace")
        locus_name, accession, version, definition = "", "", "", []
        section = None
        for line in header.splitlines():
            head = line[:12].strip() if line[:1] != " " else ""
            if head:
                section = head
                rest = line[12:].strip()
                if head == "LOCUS":
                    locus_name = rest.split()[0] if rest else ""
                elif head == "DEFINITION":
                    definition = [rest]
                elif head == "ACCESSION" and rest:
                    accession = rest.split()[0]
                elif head == "VERSION" and rest:
                    version = rest.split()[0]
            elif section == "DEFINITION":
                definition.append(line.strip())
        feat_block = ""
        if f0 >= 0:
            block = rec[f0 + 1:o0 if o0 > f0 else len(rec)].decode("ascii", "replace")
            nl = block.find("\n")
            feat_block = block[nl + 1:] if nl >= 0 else ""       # drop the FEATURES header line
            # anything after the table that starts in column 1 (CONTIG, BASE COUNT, ...) is not a
            # feature: feature lines are indented, so the table ends at the first such line
            cut = len(feat_block)
            for word in ("\nCONTIG", "\nBASE COUNT", "\nWGS", "\nCOMMENT"):
                k = feat_block.find(word)
                if 0 <= k < cut:
                    cut = k + 1
            feat_block = feat_block[:cut]
        sequence = b""
        if o0 >= 0:
            nl = rec.find(b"\n", o0 + 1)
            if nl >= 0:
                sequence = rec[nl + 1:].translate(_UPPER, b" 0123456789\t\r\n")
        description = " ".join(definition)
        if description.endswith("."):
            description = description[:-1]
        records.append(GenBankRecord(version or accession or locus_name, description, sequence,
                                     _parse_features(feat_block)))
    return records


def read(text):
    """Exactly one record, as ``SeqIO.read`` (cgb/chromid.py:34)."""
    records = parse(text)
    if len(records) != 1:
        raise ValueError("%s records found in handle" % ("No" if not records else "More than one"))
    return records[0]


def read_file(path):
    with open(path, "rb") as f:
        return read(f.read())
