"""Generate tests/golden/*.json by running the reference VERBATIM.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Executes the unmodified ``cgb/pssm_model.py``, ``cgb/binding_model.py`` and
``cgb/site_collection.py`` of the reference through oracle/ref_loader.py (import
shims + the Biopython stand-in) and records inputs and outputs.  The fixtures are
what pins oracle/np_oracle.py; the GPU box has no reference tree and reads only
the committed JSON.

Environment note recorded in every file: this container has numpy 2.x, where the
verbatim soft-max line ``2**score`` (pssm_model.py:130) is evaluated in float32
(NEP 50); the reference's pinned numpy 1.x evaluates it in float64.  The oracle
reproduces the recorded values bit-for-bit in ``softmax="nep50"`` mode and is
used against the CUDA path in ``legacy64`` mode (difference < 2e-7 per score).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_loader  # noqa: E402


def rand_dna(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))


def matrix_of(pssm):
    return {l: [float(x) for x in pssm[l]] for l in "ACGT"}


def main():
    ns = ref_loader.load()
    with open(os.path.join(ref_loader.REFERENCE_ROOT, "test_input.json")) as f:
        ti = json.load(f)
    motifs = [{"name": mo["name"], "sites": mo["sites"]} for mo in ti["motifs"]]
    meta = {"generator": "tests/golden/make_golden.py",
            "source": "reference pssm_model.py/binding_model.py/site_collection.py run verbatim",
            "numpy": np.__version__, "softmax": "nep50"}
    with open(os.path.join(HERE, "lexa_sites.json"), "w") as f:
        json.dump({"origin": "reference test_input.json:8-208 (LexA site collections; input data)",
                   "prior_regulation_probability": ti["prior_regulation_probability"],
                   "alpha": ti["alpha"],
                   "promoter_up_distance": ti["promoter_up_distance"],
                   "promoter_dw_distance": ti["promoter_dw_distance"],
                   "motifs": motifs}, f, indent=1)

    colls = ref_loader.lexa_collections(ns)
    out = {"meta": meta, "collections": [], "models": {}}

    # --- per collection: own PSSM (SURVEY §8c "survey-derived vectors")
    for c in colls:
        M = ns.PSSMModel([c], [1.0])
        site = c.sites[0]
        f = float(M.score_seq(site, both=False)[0])
        r = float(M.rev_comp_pssm.calculate(ns.pssm_model.Seq(site, "ACGT")))
        b = float(M.score_seq(site, both=True)[0])
        out["collections"].append({
            "name": c.name, "n_sites": c.site_count, "pwm": matrix_of(c.pwm),
            "IC": M.IC, "patser_threshold": M.threshold(),
            "first_site": site, "first_site_fwd": f, "first_site_rc": r, "first_site_both": b})

    # --- mixed models
    counts = [c.site_count for c in colls]
    weightings = {"equal": [1.0 / len(colls)] * len(colls),
                  "site_count": [x / float(sum(counts)) for x in counts]}
    seq2k = rand_dna(2000, 11)
    seq_n = seq2k[:700] + "N" * 7 + seq2k[707:1200] + "NRY" + seq2k[1203:]
    seq_lower = seq2k[:300] + seq2k[300:360].lower() + seq2k[360:690] + \
        seq2k[690:730].lower().replace(seq2k[700].lower(), "n", 1) + seq2k[730:]
    promoters = [rand_dna(350, 100 + i) for i in range(6)]
    for name, w in weightings.items():
        M = ns.PSSMModel(colls, w)
        rec = {"weights": w, "pwm": matrix_of(M.pwm), "pssm": matrix_of(M.pssm),
               "rev_comp_pssm": matrix_of(M.rev_comp_pssm), "IC": M.IC, "length": M.length,
               "patser_threshold": M.threshold(),
               "score_self": [float(s[0]) for s in M.score_self()]}
        cases = {}
        for cname, s in (("random2k", seq2k), ("with_N", seq_n), ("lowercase", seq_lower)):
            both = [float(x) for x in M.score_seq(s, both=True)]
            fw = [float(x) for x in M.score_seq(s, both=False)]
            cases[cname] = {"seq": s, "both": both, "fwd_f32": fw}
        # len(seq) == m quirks: scalar wrapped in a list; repaired value stays float64
        site = M.sites[3]
        site_n = site[:5] + "N" + site[6:]
        site_l = site[:9].lower() + "N" + site[10:]
        cases["single"] = {s: {"both": float(M.score_seq(s)[0]),
                               "fwd": float(M.score_seq(s, both=False)[0])}
                           for s in (site, site_n, site_l)}
        rec["score_seq"] = cases
        # Bayesian estimator fitted on all windows of seq2k, then posteriors
        bg = [[x] for x in M.score_seq(seq2k)]        # list of 1-element lists, genome.py:222-223
        M.build_bayesian_estimator(bg)
        rec["estimator"] = {"mu_m": float(M._mu_m), "std_m": float(M._std_m),
                            "mu_bg": float(M._mu_bg), "std_bg": float(M._std_bg)}
        site0 = M.sites[0]
        proms = promoters + [promoters[0][:150] + site0 + promoters[0][172:],
                             promoters[1][:40] + "NNNN" + promoters[1][44:]]
        rec["posteriors"] = [{"seq": p, "p_motif": ti["prior_regulation_probability"],
                              "alpha": ti["alpha"],
                              "posterior": float(M.binding_probability(
                                  p, ti["prior_regulation_probability"], ti["alpha"]))}
                             for p in proms]
        rec["posteriors"].append({"seq": proms[6], "p_motif": 0.01, "alpha": 1 / 350.0,
                                  "posterior": float(M.binding_probability(proms[6], 0.01))})
        try:
            M.threshold("other")
            rec["threshold_other"] = "no error"
        except ValueError:
            rec["threshold_other"] = "ValueError"
        out["models"][name] = rec

    # --- hand-derivable KAT: sites=["ACGT"], pseudocount 1
    c = ns.SiteCollection(["ACGT"], ref_loader._TF(), "kat")
    M = ns.PSSMModel([c], [1.0])
    out["kat_acgt"] = {"pwm": matrix_of(c.pwm), "pssm": matrix_of(M.pssm),
                       "fwd": float(M.score_seq("ACGT", both=False)[0]),
                       "both": float(M.score_seq("ACGT")[0])}
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(HERE, "reference_vectors.json"))


if __name__ == "__main__":
    main()
