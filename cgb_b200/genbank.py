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


def _finish_qualifier(quals, key, value):
    if key is None:
        return
    if value is None:
        v = ""
    else:
        v = value
        if v.startswith('"'):
            v = v[1:-1] if v.endswith('"') and len(v) >= 2 else v[1:]
            v = v.replace('""', '"')
        if key == "translation":
            v = v.replace(" ", "")
    quals.setdefault(key, []).append(v)


def _parse_features(lines):
    feats = []
    ftype, loc_parts, quals = None, [], None
    qkey, qval, in_quote = None, None, False

    def close_feature():
        nonlocal ftype, loc_parts, quals, qkey, qval, in_quote
        if ftype is not None:
            _finish_qualifier(quals, qkey, qval)
            start, end, strand, compound = parse_location("".join(loc_parts))
            feats.append(Feature(ftype, start, end, strand, compound, quals))
        ftype, loc_parts, quals, qkey, qval, in_quote = None, [], None, None, None, False

    for line in lines:
        key = line[5:20].strip()
        body = line[21:].rstrip()
        if key and not line[:5].strip():
            close_feature()
            ftype, loc_parts, quals = key, [body.strip()], {}
            continue
        if ftype is None:
            continue
        text = body.strip()
        if in_quote:
            qval += text if qkey == "translation" else " " + text
            in_quote = not _quote_closed(qval)
            continue
        if text.startswith("/"):
            _finish_qualifier(quals, qkey, qval)
            if "=" in text:
                qkey, qval = text[1:].split("=", 1)
                in_quote = qval.startswith('"') and not _quote_closed(qval)
            else:
                qkey, qval = text[1:], None
            continue
        if qkey is None:
            loc_parts.append(text)                  # location continued over several lines
        else:
            qval = (qval or "") + " " + text        # unquoted continuation
    close_feature()
    return feats


def parse(text):
    """All records of a GenBank flat file (``str`` or ``bytes``)."""
    if isinstance(text, bytes):
        text = text.decode("ascii", "replace")
    records = []
    lines = text.splitlines()
    i, n = 0, len(lines)
    while i < n:
        while i < n and not lines[i].startswith("LOCUS"):
            i += 1
        if i >= n:
            break
        locus_name = lines[i].split()[1] if len(lines[i].split()) > 1 else ""
        accession, version, definition = "", "", []
        feat_lines, seq_parts = [], []
        section = None
        i += 1
        while i < n and not lines[i].startswith("//"):
            line = lines[i]
            head = line[:12].strip() if line[:1] != " " else ""
            if head:
                section = head
                rest = line[12:].strip()
                if head == "DEFINITION":
                    definition = [rest]
                elif head == "ACCESSION" and rest:
                    accession = rest.split()[0]
                elif head == "VERSION" and rest:
                    version = rest.split()[0]
                elif head.startswith("FEATURES"):
                    section = "FEATURES"
                elif head.startswith("ORIGIN"):
                    section = "ORIGIN"
            elif section == "DEFINITION":
                definition.append(line.strip())
            elif section == "FEATURES":
                feat_lines.append(line)
            elif section == "ORIGIN":
                seq_parts.append(line)
            i += 1
        i += 1
        raw = "".join(seq_parts).encode("ascii")
        sequence = raw.translate(_UPPER, b" 0123456789\t\r\n")
        description = " ".join(definition)
        if description.endswith("."):
            description = description[:-1]
        records.append(GenBankRecord(version or accession or locus_name, description, sequence,
                                     _parse_features(feat_lines)))
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
