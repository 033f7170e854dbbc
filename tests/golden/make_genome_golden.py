"""Generate tests/golden/genome_vectors.json by running the reference's genome-side stack VERBATIM.

Run in the build container only (needs /root/reference):

    python tests/golden/make_genome_golden.py

``cgb/genome.py``, ``cgb/chromid.py``, ``cgb/gene.py``, ``cgb/operon.py``, ``cgb/bio_utils.py`` (and
the model files) are executed unmodified through oracle/ref_loader.py on a small synthetic
annotated genome (two replicons; the feature tables are built here and handed to the reference as
the SeqRecord its GenBank parser would have produced).  Recorded: the gene table the reference
derives (Chromid.genes, chromid.py:99-122 -- incl. the skipped compound-location genes, the
gene/product pairing and the dropped last feature), the coordinate rules (gene.py:57-127), the
posterior of regulation of every gene (genome.py:330-333), the identified sites and their CSV
(genome.py:394-494), operon prediction with and without posterior-driven splitting
(chromid.py:173-243, genome.py:74-116) and the operon / regulon CSVs (genome.py:130-152, 335-378).

The fixture holds the inputs as well (sequences + feature tables), so the tests need nothing else.
"""
from __future__ import annotations

import json
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def make_replicon(accession, length, seed, planted_site, end_with_gene=False):
    """Sequence + GenBank-ordered feature table (plain dicts) of one synthetic replicon."""
    rng = np.random.Generator(np.random.PCG64(seed))
    seq = np.array(list("ACGT"))[rng.integers(0, 4, size=length)]
    feats = [{"type": "source", "start": 0, "end": length, "strand": 1, "compound": False, "qualifiers": {}}]
    pos, strand, k = int(rng.integers(150, 400)), 1, 0
    genes = []
    while True:
        glen = int(rng.integers(300, 1500))
        if pos + glen > length - 120:
            break
        k += 1
        tag = "%s_%04d" % (accession[:3], 10 * k)
        kind = rng.choice(["cds", "cds_noid", "trna", "none", "misc", "compound"],
                          p=[0.70, 0.05, 0.08, 0.07, 0.05, 0.05])
        gene_q = {"locus_tag": [tag]}
        if rng.random() < 0.4:
            gene_q["gene"] = ["gen%d" % k]
        g = {"type": "gene", "start": pos, "end": pos + glen, "strand": strand,
             "compound": kind == "compound", "qualifiers": gene_q}
        feats.append(g)
        genes.append(g)
        if kind in ("cds", "cds_noid", "compound"):
            q = {"locus_tag": [tag], "product": ["protein %d, putative" % k]}   # a comma: CSV quoting
            if kind != "cds_noid":
                q["protein_id"] = ["WP_%09d.1" % (seed * 1000 + k)]
            feats.append({"type": "CDS", "start": pos, "end": pos + glen, "strand": strand,
                          "compound": kind == "compound", "qualifiers": q})
        elif kind == "trna":
            feats.append({"type": "tRNA", "start": pos, "end": pos + glen, "strand": strand, "compound": False,
                          "qualifiers": {"locus_tag": [tag], "product": ["tRNA-Ala"]}})
        elif kind == "misc":
            feats.append({"type": "misc_feature", "start": pos + 10, "end": pos + 50, "strand": strand,
                          "compound": False, "qualifiers": {"note": ["other"]}})
        # next gene: operon-like short gap (may overlap) or a long one, where the strand may flip
        if rng.random() < 0.55:
            gap = int(rng.integers(-12, 45))
        else:
            gap = int(rng.integers(80, 600))
            if rng.random() < 0.45:
                strand = -strand
        pos = pos + glen + gap
    if end_with_gene:
        # the reference pairs each feature with its successor (chromid.py:103): a gene that is
        # the very last feature is never seen
        tag = "%s_LAST" % accession[:3]
        feats.append({"type": "gene", "start": length - 110, "end": length - 20, "strand": 1, "compound": False,
                      "qualifiers": {"locus_tag": [tag]}})
    # strong sites upstream of a few genes, an N run and a lower-case stretch in intergenic space
    plain = [g for g in genes if not g["compound"]]
    for j, g in enumerate(plain[3::9]):
        site = planted_site if j % 2 == 0 else ref_loader_revcomp(planted_site)
        at = g["start"] - 40 - len(site) if g["strand"] == 1 else g["end"] + 40
        if 0 <= at and at + len(site) <= length:
            seq[at:at + len(site)] = list(site)
    g = plain[5]
    seq[max(0, g["start"] - 30):max(0, g["start"] - 30) + 9] = "N"
    g = plain[8]
    lo = max(0, g["start"] - 120)
    seq[lo:lo + 60] = [c.lower() for c in seq[lo:lo + 60]]
    return {"accession": accession, "description": "synthetic replicon %s" % accession,
            "sequence": "".join(seq), "features": feats}


def ref_loader_revcomp(s):
    return s.translate(str.maketrans("ACGT", "TGCA"))[::-1]


def to_record(rep):
    feats = []
    for f in rep["features"]:
        if f["compound"]:
            mid = (f["start"] + f["end"]) // 2
            loc = ref_loader.CompoundLocation([(f["start"], mid - 3), (mid + 3, f["end"])], f["strand"])
        else:
            loc = ref_loader.FeatureLocation(f["start"], f["end"], f["strand"])
        feats.append(ref_loader.FakeFeature(f["type"], loc, {k: list(v) for k, v in f["qualifiers"].items()}))
    return ref_loader.FakeRecord(rep["accession"], rep["description"], rep["sequence"], feats)


def read_text(path):
    with open(path, newline="") as f:
        return f.read()


def main():
    ns = ref_loader.load_genome_stack()
    with open(os.path.join(ref_loader.REFERENCE_ROOT, "test_input.json")) as f:
        ti = json.load(f)
    colls = ref_loader.lexa_collections(ns)
    counts = [c.site_count for c in colls]
    weights = [x / float(sum(counts)) for x in counts]
    planted = colls[0].sites[0]
    reps = [make_replicon("SYNA01.1", 80000, 21, planted),
            make_replicon("SYNB02.1", 12000, 22, planted, end_with_gene=True)]
    for rep in reps:
        ref_loader.register_record(to_record(rep))
    genome = ns.Genome("synthetic_strain", [rep["accession"] for rep in reps])

    out = {"meta": {"generator": "tests/golden/make_genome_golden.py",
                    "source": "reference genome.py/chromid.py/gene.py/operon.py/bio_utils.py run verbatim",
                    "numpy": np.__version__, "softmax": "nep50"},
           "weights": weights, "prior": ti["prior_regulation_probability"], "alpha": ti["alpha"],
           "promoter_up_distance": ti["promoter_up_distance"], "promoter_dw_distance": ti["promoter_dw_distance"],
           "replicons": reps, "genes": [], "scenarios": {}}

    # --- the gene table the reference derives from the feature tables + coordinate rules
    up, dw = ti["promoter_up_distance"], ti["promoter_dw_distance"]
    for c in genome.chromids:
        rows = []
        for g in c.genes:
            try:
                ug = g.upstream_gene
                ug_tag = ug.locus_tag if ug is not None else None
                loc_none = list(g.upstream_noncoding_region_location(None, dw))
            except IndexError:                  # genes[_index - 1] past the shortened list
                ug_tag, loc_none = "IndexError", None
            rows.append({"index": g._index, "start": g.start, "end": g.end, "strand": g.strand,
                         "locus_tag": g.locus_tag, "name": g.name, "product_type": g.product_type,
                         "product": g.product, "is_protein_coding": g.is_protein_coding_gene,
                         "protein_id": g.protein_accession_number if g.is_protein_coding_gene else "",
                         "upstream_gene": ug_tag,
                         "upstream_noncoding_none": loc_none,
                         "upstream_noncoding_up": list(g.upstream_noncoding_region_location(up, dw)),
                         "promoter_region": g.promoter_region(up, dw)})
        out["genes"].append({"accession": c.accession_number, "length": c.length, "rows": rows})

    # --- model: the reference's own Genome.build_PSSM_model (genome.py:203-227) run VERBATIM, with
    # random.sample made the identity so that the background is the whole population the sample
    # is drawn from -- every m-mer (last one dropped, genome.py:218) of every gene's upstream
    # non-coding region, each scored by score_seq as a sequence of its own -- and the result is
    # reproducible.  (The reference scores a random 10 000 of them.)
    sampled = {}

    def identity_sample(population, k):
        sampled["population"], sampled["k"] = len(population), k
        return list(population)

    real_random = ns.genome.random
    ns.genome.random = types.SimpleNamespace(sample=identity_sample)
    try:
        genome.build_PSSM_model(colls, weights)
    finally:
        ns.genome.random = real_random
    model = genome.TF_binding_model
    out["estimator"] = {"mu_m": float(model._mu_m), "std_m": float(model._std_m),
                        "mu_bg": float(model._mu_bg), "std_bg": float(model._std_bg),
                        "threshold": float(model.threshold()), "n_bg": sampled["population"],
                        "population": "upstream non-coding m-mers (genome.py:215-218), random.sample = identity",
                        "reference_sample_size": sampled["k"]}
    # the same model with the background over ALL windows of every replicon (the opt-in
    # mode="all" of the new build; north_star's genome-wide distribution)
    model_all = ns.PSSMModel(colls, weights)
    bg_scores = []
    for c in genome.chromids:
        bg_scores.extend(model_all.score_seq(c.sequence))
    model_all.build_bayesian_estimator(bg_scores)
    out["estimator_all_windows"] = {"mu_bg": float(model_all._mu_bg), "std_bg": float(model_all._std_bg),
                                    "n_bg": len(bg_scores)}

    tmp = tempfile.mkdtemp()
    for scen, use_up in (("scan_to_upstream_gene", False), ("scan_up_distance", True)):
        ui = types.SimpleNamespace(promoter_up_distance=up, promoter_dw_distance=dw, alpha=ti["alpha"],
                                   use_up_dist_site_scan=use_up)
        # genes whose upstream gene lookup raises IndexError would abort identify_sites: the
        # synthetic annotation keeps the compound genes away from that case (asserted here)
        path = os.path.join(tmp, "sites_%s.csv" % scen)
        genome.identify_sites(ui, path)
        sites = [{"chromid": s.chromid.accession_number, "start": int(s.start), "end": int(s.end),
                  "strand": int(s.strand), "score": float(s.score), "gene": s.gene.locus_tag}
                 for s in genome.putative_sites]
        out["scenarios"][scen] = {"sites": sites, "sites_csv": read_text(path)}

    ui = types.SimpleNamespace(promoter_up_distance=up, promoter_dw_distance=dw, alpha=ti["alpha"],
                               use_up_dist_site_scan=False)
    genome.calculate_regulation_probabilities(ti["prior_regulation_probability"], ui)
    out["regulation_probability"] = [[float(g.regulation_probability) for g in c.genes] for c in genome.chromids]

    # --- operons: with posterior-driven splitting (th 0.5) and without (th 1.0), sigma 1.0 / 1.5
    out["operons"] = {}
    for name, th, sigma in (("split_0.5_sigma_1.0", 0.5, 1.0), ("nosplit_sigma_1.0", 1.0, 1.0),
                            ("split_0.2_sigma_1.5", 0.2, 1.5)):
        genome.operon_prediction(th, sigma)
        opr_csv = os.path.join(tmp, "operons_%s.csv" % name)
        genome.operons_to_csv(opr_csv)
        reg_csv = os.path.join(tmp, "regulons_%s.csv" % name)
        regulons = genome.infer_regulons(0.5, reg_csv)
        out["operons"][name] = {
            "probability_th": th, "sigma": sigma,
            "distance_threshold": float(genome.intergenic_distance_threshold(sigma)),
            "operons": [{"id": o.operon_id, "chromid": o.chromid.accession_number, "start": o.start, "end": o.end,
                         "strand": o.strand, "genes": [g.locus_tag for g in o.genes],
                         "first_gene": o.first_gene.locus_tag,
                         "regulation_probability": float(o.regulation_probability)} for o in genome.operons],
            "operons_csv": read_text(opr_csv), "regulons_csv": read_text(reg_csv),
            "regulons": [[g.locus_tag for g in o.genes] for o in regulons]}
        genome.remove_operons() if False else None      # the reference's remove_operons mutates a temp list

    # --- get_prior (cgb/__init__.py:320-350; that file is Python-2-only syntax, so its three
    # lines are spelled out here over the verbatim Genome): operons without splitting, then
    # G / 2^IC expected sites over the number of operons
    genome.operon_prediction(1, 1.0)
    out["inferred_prior"] = {"sigma": 1.0, "IC": float(model.IC), "genome_length": genome.length,
                             "num_operons": genome.num_operons,
                             "prior": genome.length / 2 ** genome.TF_binding_model.IC / genome.num_operons}

    with open(os.path.join(HERE, "genome_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    n_genes = sum(len(c["rows"]) for c in out["genes"])
    print("genes", n_genes, "sites", {k: len(v["sites"]) for k, v in out["scenarios"].items()},
          "operons", {k: len(v["operons"]) for k, v in out["operons"].items()},
          "max P", max(max(p) for p in out["regulation_probability"]))


if __name__ == "__main__":
    random.seed(7)
    main()
