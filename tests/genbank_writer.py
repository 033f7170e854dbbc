"""Test helper: writes a GenBank flat file (the layout NCBI serves and Biopython's SeqIO 'gb' parser
reads) from a replicon spec of tests/golden/genome_vectors.json."""


def _location(f):
    if f["compound"]:
        mid = (f["start"] + f["end"]) // 2
        loc = "join(%d..%d,%d..%d)" % (f["start"] + 1, mid - 3, mid + 4, f["end"])
    else:
        loc = "%d..%d" % (f["start"] + 1, f["end"])
    return "complement(%s)" % loc if f["strand"] == -1 else loc


def _wrap(text, width=58):
    out = []
    while len(text) > width:
        cut = text.rfind(" ", 0, width + 1)
        if cut <= 0:
            cut = width
        out.append(text[:cut])
        text = text[cut:].lstrip(" ")
    out.append(text)
    return out


def genbank_text(rep, fuzzy_first=True):
    acc, ver = rep["accession"].split(".")[0], rep["accession"]
    n = len(rep["sequence"])
    lines = ["LOCUS       %-16s %11d bp    DNA     circular BCT 01-JAN-2020" % (acc, n),
             "DEFINITION  %s." % rep["description"],
             "ACCESSION   %s" % acc,
             "VERSION     %s" % ver,
             "KEYWORDS    .",
             "SOURCE      synthetic construct",
             "  ORGANISM  synthetic construct",
             "            other sequences; artificial sequences.",
             "FEATURES             Location/Qualifiers"]
    first_gene = True
    for f in rep["features"]:
        loc = _location(f)
        if fuzzy_first and f["type"] == "gene" and first_gene and not f["compound"] and f["strand"] == 1:
            loc = "<" + loc                  # fuzzy start: Biopython's .position drops the '<'
            first_gene = False
        lines.append("     %-15s %s" % (f["type"], loc))
        if f["type"] == "source":
            lines.append('                     /organism="synthetic construct"')
            lines.append('                     /mol_type="genomic DNA"')
        for key, vals in f["qualifiers"].items():
            for v in vals:
                text = '/%s="%s"' % (key, v.replace('"', '""'))
                if key == "product":
                    text = '/%s="%s"' % (key, v + " " * 0)
                for part in _wrap(text):
                    lines.append("                     " + part)
        if f["type"] == "CDS":
            lines.append("                     /codon_start=1")
            lines.append("                     /transl_table=11")
            lines.append('                     /translation="MKKLLPIAVAGLLSACSHEELAAKLNQYQVRRDAIRERMA')
            lines.append('                     QLEDEIRRLKEALR"')
        if f["type"] == "gene" and f["strand"] == -1:
            lines.append("                     /pseudo")
    lines.append("ORIGIN")
    seq = rep["sequence"].lower()
    for i in range(0, n, 60):
        chunk = seq[i:i + 60]
        lines.append("%9d %s" % (i + 1, " ".join(chunk[j:j + 10] for j in range(0, len(chunk), 10))))
    lines.append("//")
    return "\n".join(lines) + "\n"
