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


class FeatureIndex(object):
    """Structural index of a feature table built by the library's host code
    (``cgbk_genbank_index``: one C pass instead of a Python walk over every line): per feature the
    key, the parsed location and its qualifier range; per qualifier the key and the RAW value span.
    Values are decoded on demand -- the gene table needs five qualifiers out of dozens."""

    def __init__(self, block):
        import ctypes as C

        import numpy as np

        from . import _cabi
        self.block = block                                   # bytes of the table (no FEATURES header line)
        lines = block.count(b"\n") + 2
        feat = np.empty((lines, 10), dtype=np.int32)
        qual = np.empty((lines, 5), dtype=np.int32)
        nf, nq = C.c_int32(), C.c_int32()
        _cabi.check(_cabi.lib().cgbk_genbank_index(block, len(block), feat.ctypes.data, lines, qual.ctypes.data,
                                                   lines, C.byref(nf), C.byref(nq)))
        self.feat = feat[:nf.value]
        self.qual = qual[:nq.value]
        self._types = None

    def __len__(self):
        return len(self.feat)

    @property
    def types(self):
        if self._types is None:
            b = self.block
            self._types = [b[o:o + n].decode("ascii", "replace") for o, n in self.feat[:, :2].tolist()]
        return self._types

    def qualifier_values(self, k, key):
        """Decoded values of qualifier ``key`` (bytes) of feature ``k``, in file order; ``None`` when
        the feature has no such qualifier."""
        b = self.block
        q0, qn = int(self.feat[k, 4]), int(self.feat[k, 5])
        out = None
        for ko, kn, vo, vn, _ in self.qual[q0:q0 + qn].tolist():
            if kn == len(key) and b[ko:ko + kn] == key:
                if out is None:
                    out = []
                out.append("" if vn < 0 else _clean_value(key.decode("ascii"), b[vo:vo + vn].decode("ascii", "replace")))
        return out

    def feature(self, k):
        b = self.block
        f = self.feat[k].tolist()
        quals = {}
        for ko, kn, vo, vn, _ in self.qual[f[4]:f[4] + f[5]].tolist():
            key = b[ko:ko + kn].decode("ascii", "replace")
            if vn < 0:
                quals.setdefault(key.strip(), []).append("")
            else:
                quals.setdefault(key, []).append(_clean_value(key, b[vo:vo + vn].decode("ascii", "replace")))
        return Feature(self.types[k], f[6], f[7], f[8], bool(f[9]), quals)


class GenBankRecord(object):
    """One LOCUS ... // entry.  ``features`` may be given as a list of :class:`Feature` or as a
    :class:`FeatureIndex` (then the Feature tuples are only built when somebody asks for them)."""

    def __init__(self, accession, description, sequence, features):
        self.id = accession
        self.description = description
        self.sequence = sequence            # bytes, upper case
        self._index = features if isinstance(features, FeatureIndex) else None
        self._features = None if self._index is not None else features

    @property
    def features(self):
        if self._features is None:
            self._features = [self._index.feature(k) for k in range(len(self._index))]
        return self._features

    @property
    def length(self):
        return len(self.sequence)

    def gene_table(self, seq_index=0):
        """``Chromid.genes`` (cgb/chromid.py:99-122) as a GeneTable with the annotation columns
        the reports use (locus_tag, gene name, product type, product, protein_id)."""
        if self._index is not None and self._features is None:
            return self._gene_table_from_index(seq_index)
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


    def _gene_table_from_index(self, seq_index):
        """The same selection over the feature index: the qualifier keys are matched with numpy
        over the whole table, and only the values of ``gene`` features and of the feature after
        each are decoded."""
        import numpy as np
        ix = self._index
        feat, qual, b = ix.feat, ix.qual, ix.block
        nf = len(feat)
        raw = np.frombuffer(b, dtype=np.uint8)

        def spans_equal(off, length, word):
            """mask over spans: the span text equals ``word`` (bytes)."""
            m = length == len(word)
            idx = np.nonzero(m)[0]
            if len(idx):
                cols = off[idx][:, None].astype(np.int64) + np.arange(len(word))
                m[idx] = (raw[cols] == np.frombuffer(word, dtype=np.uint8)).all(axis=1)
            return m

        is_gene = spans_equal(feat[:, 0], feat[:, 1], b"gene")
        feature_of_qual = np.repeat(np.arange(nf), feat[:, 5])

        class ByFeature(object):
            """feature -> [qualifier rows] of one key, file order (first row and count per feature as
            arrays; the rare repeated key is collected on demand)."""

            def __init__(self, word):
                rows = np.nonzero(spans_equal(qual[:, 0], qual[:, 1], word))[0]
                f = feature_of_qual[rows]
                self.count = np.bincount(f, minlength=nf)
                self.first = np.full(nf, -1, dtype=np.int64)
                self.first[f[::-1]] = rows[::-1]                 # the earliest row wins
                self.rows, self.f = rows, f

            def get(self, i):
                c = self.count[i]
                if c == 0:
                    return None
                if c == 1:
                    return [int(self.first[i])]
                return self.rows[self.f == i].tolist()

        q_locus, q_gene, q_product, q_protein = (ByFeature(w) for w in (b"locus_tag", b"gene", b"product",
                                                                       b"protein_id"))
        qv = qual[:, 2:4].tolist()
        simple = qual[:, 4].tolist()

        def value(q, key):
            vo, vn = qv[q]
            if vn < 0:
                return ""
            r = b[vo:vo + vn]
            if simple[q]:                               # one line, quoted, no doubled quote: nothing to clean
                return r[1:-1].decode("ascii", "replace")
            return _clean_value(key, r.decode("ascii", "replace"))

        genes = np.nonzero(is_gene[:nf - 1])[0]                    # the last feature is never looked at
        fast = self._gene_table_simple(seq_index, genes, feat, qual, raw, b, q_locus, q_gene, q_product, q_protein)
        if fast is not None:
            return fast
        start, end, strand, index = [], [], [], []
        locus, name, ptype, product, protein = [], [], [], [], []
        k = 0
        loc = feat[:, 6:10].tolist()
        for i in genes.tolist():
            qs = q_locus.get(i)
            if qs is None:
                raise KeyError("locus_tag")                       # as in the reference
            tags = [value(q, "locus_tag") for q in qs]
            nq = q_locus.get(i + 1)
            follows = nq is not None and tags == [value(q, "locus_tag") for q in nq]
            st, en, sd, compound = loc[i]
            if compound:
                k += 1
                continue
            assert len(tags) == 1                                 # Gene.locus_tag, gene.py:178
            start.append(st); end.append(en); strand.append(sd); index.append(k)
            locus.append(tags[0])
            g = q_gene.get(i)
            name.append(value(g[0], "gene") if g is not None else tags[0])
            if follows:
                o, n = feat[i + 1, 0], feat[i + 1, 1]
                pt = b[o:o + n].decode("ascii", "replace")
                pr = q_product.get(i + 1)
                product.append(value(pr[0], "product") if pr is not None else "")
                pid = q_protein.get(i + 1) if pt == "CDS" else None
                protein.append(value(pid[0], "protein_id") if pid is not None else "")
            else:
                pt = ""
                product.append("")
                protein.append("")
            ptype.append(pt)
            k += 1
        return GeneTable(start, end, strand, self.length, seq_index=seq_index, locus_tag=locus,
                         protein_id=protein, product=product, index=index, name=name, product_type=ptype)


    def _gene_table_simple(self, seq_index, genes, feat, qual, raw, b, q_locus, q_gene, q_product, q_protein):
        """The whole table with array operations, for the layout every annotation pipeline writes:
        each needed qualifier at most once per feature, its value quoted on one line with no
        doubled quote.  Returns None when any needed value is not that simple (the caller then
        walks the genes one by one)."""
        import numpy as np
        if len(genes) == 0:
            return None
        nxt = genes + 1
        cl = q_locus.count
        if (cl[genes] != 1).any() or (cl[nxt] > 1).any():
            return None
        first = {"l": q_locus.first, "g": q_gene.first, "p": q_product.first, "i": q_protein.first}
        for cnt, idx in ((q_gene.count, genes), (q_product.count, nxt), (q_protein.count, nxt)):
            if (cnt[idx] > 1).any():
                return None
        rows = np.concatenate([first["l"][genes], first["l"][nxt], first["g"][genes], first["p"][nxt], first["i"][nxt]])
        rows = rows[rows >= 0]
        if not qual[rows, 4].all():                   # the indexer marks one-line quoted values without inner quotes
            return None

        def strings(q):
            """decoded values of qualifier rows ``q`` (-1: no such qualifier -> None)"""
            off = (qual[q, 2].astype(np.int64) + 1).tolist()
            ln = (qual[q, 3].astype(np.int64) - 2).tolist()
            out = b"\x00".join([b[o:o + n] for o, n in zip(off, ln)]).decode("ascii", "replace").split("\x00")
            return [v if r >= 0 else None for v, r in zip(out, q.tolist())]

        tags = strings(first["l"][genes])
        tags_n = strings(first["l"][nxt])
        follows = np.array([a == c for a, c in zip(tags, tags_n)], dtype=bool)
        keep = feat[genes, 9] == 0                                    # one contiguous span
        types_n = [b[o:o + n].decode("ascii", "replace") for o, n in feat[nxt, :2].tolist()]
        gname = strings(first["g"][genes])
        prod = strings(first["p"][nxt])
        pid = strings(first["i"][nxt])
        sel = np.nonzero(keep)[0].tolist()
        fl = follows.tolist()
        ptype = [types_n[j] if fl[j] else "" for j in sel]
        return GeneTable(feat[genes, 6][keep].tolist(), feat[genes, 7][keep].tolist(), feat[genes, 8][keep].tolist(),
                         self.length, seq_index=seq_index, locus_tag=[tags[j] for j in sel],
                         protein_id=[(pid[j] or "") if fl[j] and types_n[j] == "CDS" else "" for j in sel],
                         product=[(prod[j] or "") if fl[j] else "" for j in sel], index=sel,
                         name=[gname[j] if gname[j] is not None else tags[j] for j in sel], product_type=ptype)


def _clean_value(key, value):
    """Raw qualifier value (continuation lines and quotes still in it) -> the string Biopython
    reports: lines joined (with a blank; /translation without), outer quotes off, "" -> "."""
    if "\n" in value:
        value = _CONTINUATION.sub("" if key == "translation" else " ", value)
    value = value.rstrip()
    if value.startswith('"'):
        value = value[1:-1] if value.endswith('"') and len(value) >= 2 else value[1:]
        if '""' in value:
            value = value.replace('""', '"')
    if key == "translation" and " " in value:
        value = value.replace(" ", "")
    return value


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


def _index_or_parse(feat_block):
    """The C index of the table (libcgbk host code, no GPU needed); the pure-Python walk only when
    the library has not been built (the parser is host logic and must stay importable then)."""
    try:
        return FeatureIndex(feat_block)
    except Exception as exc:                                  # library missing / too old
        from . import _cabi
        if isinstance(exc, _cabi.CgbkError) or isinstance(exc, (OSError, AttributeError)):
            # anything after the table that starts in column 1 (CONTIG, BASE COUNT, ...) is not a
            # feature: feature lines are indented, so the table ends at the first such line (the C
            # index stops there by itself)
            cut = len(feat_block)
            for word in (b"\nCONTIG", b"\nBASE COUNT", b"\nWGS", b"\nCOMMENT"):
                k = feat_block.find(word)
                if 0 <= k < cut:
                    cut = k + 1
            return _parse_features(feat_block[:cut].decode("ascii", "replace"))
        raise


def _origin_sequence(text, first_line):
    """Sequence bytes of the ORIGIN block whose first data line starts at ``first_line`` and the
    offset of the record's closing ``//`` line: one C pass (``cgbk_genbank_origin``); without the
    library a ``find`` and one ``bytes.translate``."""
    try:
        import ctypes as C

        import numpy as np

        from . import _cabi
        lib = _cabi.lib()
        out = np.empty(max(len(text) - first_line, 1), dtype=np.uint8)
        nb, end = C.c_int64(), C.c_int64()
        _cabi.check(lib.cgbk_genbank_origin(text, len(text), first_line, out.ctypes.data, out.size,
                                            C.byref(nb), C.byref(end)))
        return out[:nb.value].tobytes(), end.value
    except (OSError, AttributeError):                          # library missing / too old
        end = first_line if text.startswith(b"//", first_line) else text.find(b"\n//", first_line)
        end = len(text) if end < 0 else end + (0 if text.startswith(b"//", first_line) else 1)
        return text[first_line:end].translate(_UPPER, b" 0123456789\t\r\n"), end


def parse(text):
    """All records of a GenBank flat file (``str`` or ``bytes``).  The blocks of a record (header,
    feature table, ORIGIN) are located with ``find``; the feature table is indexed and the ORIGIN
    block decoded by the library's host code, one C pass each."""
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
        # the record's blocks, in file order: header | FEATURES | ORIGIN | //  (a record without
        # ORIGIN ends at its "//" line)
        f0 = text.find(b"\nFEATURES", start)
        o0 = text.find(b"\nORIGIN", f0 if f0 >= 0 else start)
        if o0 < 0 or text.find(b"\n//", f0 if f0 >= 0 else start, o0) >= 0 or \
                (f0 >= 0 and text.find(b"\nLOCUS", start + 1, f0) >= 0):
            # no ORIGIN (or FEATURES) in THIS record: bound it by its own terminator first
            end = text.find(b"\n//", start)
            end = n if end < 0 else end
            f0 = text.find(b"\nFEATURES", start, end)
            o0 = text.find(b"\nORIGIN", start, end)
        else:
            end = None
        if f0 >= 0 and o0 >= 0 and o0 < f0:
            o0 = -1
        head_end = f0 if f0 >= 0 else (o0 if o0 >= 0 else (end if end is not None else n))
        header = text[start:head_end].decode("ascii", "replace")
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
        sequence = b""
        if o0 >= 0:
            nl = text.find(b"\n", o0 + 1)
            if nl >= 0:
                sequence, end_seq = _origin_sequence(text, nl + 1)
                if end is None:
                    end = end_seq
        if end is None:
            end = text.find(b"\n//", start)
            end = n if end < 0 else end
        feat_block = b""
        if f0 >= 0:
            block_end = o0 if o0 > f0 else end
            nl = text.find(b"\n", f0 + 1, block_end)
            feat_block = text[nl + 1:block_end] if nl >= 0 else b""      # drop the FEATURES header line
        pos = end + 2
        description = " ".join(definition)
        if description.endswith("."):
            description = description[:-1]
        records.append(GenBankRecord(version or accession or locus_name, description, sequence,
                                     _index_or_parse(feat_block)))
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
