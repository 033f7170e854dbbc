"""Host-side mirror of the reference's ``Genome`` for the accelerated path (cgb/genome.py).

Same method names and argument meaning as the reference for the calls ``go()`` makes on the hot
path (cgb/__init__.py:686-712): ``build_PSSM_model``, ``identify_sites``,
``calculate_regulation_probabilities``, ``operon_prediction``, ``infer_regulons``,
``operons_to_csv``.  A genome is a list of replicons (sequence + GeneTable); all replicons live
in ONE packed device buffer and every method is one or two kernel launches plus array work on
the host -- there are no per-gene Python objects.

Outside the path (and absent here): Entrez download, BLAST, phylogeny, plots.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

from . import genbank, operons as _operons
from .genome_buffer import PackedGenome
from .pssm_model import PSSMModel
from .site_search import (calculate_regulation_probabilities, identified_sites_rows, identify_sites)

Chromid = namedtuple("Chromid", "accession_number description sequence gene_table")


class Genome(object):
    def __init__(self, strain_name, chromids, device=None):
        """``chromids``: list of :class:`Chromid` (``sequence`` as bytes/str; ``gene_table.seq_index``
        is set to the position in the list)."""
        self._strain_name = strain_name
        self._chromids = list(chromids)
        for k, c in enumerate(self._chromids):
            c.gene_table.seq_index = k
        self._device = device
        self._packed = None
        self._TF_binding_model = None
        self._putative_sites = None
        self._regulation_probability = None
        self._operons = []

    @classmethod
    def from_genbank(cls, strain_name, texts, device=None):
        """One GenBank flat-file text (str/bytes) per replicon, as the records
        ``Chromid.__init__`` downloads (cgb/chromid.py:32-35)."""
        chromids = []
        for k, text in enumerate(texts):
            rec = genbank.read(text)
            chromids.append(Chromid(rec.id, rec.description, rec.sequence, rec.gene_table(k)))
        return cls(strain_name, chromids, device)

    # -- plain accessors (genome.py:52-72,120-165)
    @property
    def strain_name(self):
        return self._strain_name

    @property
    def chromids(self):
        return self._chromids

    @property
    def num_chromids(self):
        return len(self._chromids)

    @property
    def length(self):
        return sum(len(c.sequence) for c in self._chromids)

    @property
    def gene_tables(self):
        return [c.gene_table for c in self._chromids]

    @property
    def packed(self):
        """All replicons in one packed device buffer (built on first use)."""
        if self._packed is None:
            self._packed = PackedGenome.from_sequences([c.sequence for c in self._chromids], self._device)
        return self._packed

    @property
    def TF_binding_model(self):
        return self._TF_binding_model

    # -- genome.py:203-227
    def upstream_regions(self, down=50):
        """Region table of the reference's background population (genome.py:215): every gene's
        ``upstream_noncoding_region_location(None, down)`` (gene.py:72-104), per replicon."""
        return [(gt.seq_index,) + tuple(gt.upstream_noncoding_region_location(None, down))
                for gt in self.gene_tables if len(gt)]

    def build_PSSM_model(self, collections, weights, mode="upstream"):
        """PSSM model + Bayesian estimator (genome.py:203-227).

        ``mode="upstream"`` (default): the background is the reference's own population -- every
        m-mer of every gene's upstream non-coding region (to the previous gene, 50 bp into the
        gene; the last m-mer of each region dropped as ``xrange(len(seq) - m)`` does), each scored
        by ``score_seq`` as a sequence of its own, regions overlapping or repeated as the gene
        table has them.  The reference then keeps a ``random.sample`` of <= 10 000 of those
        m-mers; here ALL of them are scored (deterministic; the sample's mean / std scatter
        around these values).
        ``mode="all"``: every window of every replicon, coding sequence included (the genome-wide
        distribution of the benchmark; a different population: intergenic and coding
        composition differ on real genomes, so mu_bg / std_bg shift)."""
        model = PSSMModel(collections, weights, device=self._device)
        if mode == "all":
            model.fit_background(self.packed)
        elif mode == "upstream":
            # the genome-wide scan still runs: its score plane serves the promoter posteriors;
            # the estimator's moments come from the region table
            bg = model.background_stats(self.packed, keep_scores=True)
            reg = model.background_stats(self.packed, regions=self.upstream_regions(), n_bins=bg["n_bins"],
                                         hist_lo=bg["hist_lo"], hist_hi=bg["hist_hi"])
            bg = dict(bg, stats=reg["stats"], hist=reg["hist"], work=None, work_key=None)
            model.fit_background(bg=bg)
        else:
            raise ValueError("mode must be 'upstream' or 'all'")
        self._TF_binding_model = model

    # -- genome.py:394-455
    def identify_sites(self, user_input, filename=None):
        self._putative_sites = identify_sites(
            self._TF_binding_model, self.packed, self.gene_tables, user_input.promoter_up_distance,
            user_input.promoter_dw_distance, user_input.use_up_dist_site_scan)
        if filename:
            _operons.write_rows(filename, self.identified_sites_rows(user_input.promoter_up_distance))

    @property
    def putative_sites(self):
        return self._putative_sites

    def identified_sites_rows(self, promoter_up):
        """Header + rows of Genome._output_identified_sites (genome.py:457-494)."""
        from .site_search import CSV_HEADER
        seqs = [c.sequence.decode("ascii") if isinstance(c.sequence, (bytes, bytearray)) else str(c.sequence)
                for c in self._chromids]
        return [CSV_HEADER] + identified_sites_rows(self._putative_sites, seqs, self.gene_tables, promoter_up,
                                                    [c.accession_number for c in self._chromids])

    # -- genome.py:330-333
    def calculate_regulation_probabilities(self, prior, user_input):
        """Posterior of regulation of every gene; kept per replicon as float64 arrays in gene order."""
        out = []
        for gt in self.gene_tables:
            if len(gt) == 0:
                out.append(np.zeros(0))
                continue
            p = calculate_regulation_probabilities(self._TF_binding_model, self.packed, gt, prior,
                                                   user_input.alpha, user_input.promoter_up_distance,
                                                   user_input.promoter_dw_distance)
            out.append(p.cpu().numpy())
        self._regulation_probability = out
        return out

    @property
    def regulation_probability(self):
        return self._regulation_probability

    # -- genome.py:74-116, chromid.py:173-243
    def intergenic_distance_threshold(self, sigma=1.0):
        return _operons.intergenic_distance_threshold(self.gene_tables, sigma)

    def operon_prediction(self, probability_th, sigma):
        probs = self._regulation_probability
        if probs is None:                      # fine as long as splitting is off (get_prior)
            if probability_th != 1.0:
                raise AttributeError("calculate_regulation_probabilities has not been run "
                                     "(Gene._regulation_probability, cgb/gene.py:144)")
            probs = [None] * len(self._chromids)
        self._operons = _operons.predict_operons_genome(
            self.gene_tables, probability_th, sigma, probs,
            [c.accession_number for c in self._chromids])

    @property
    def operons(self):
        return self._operons

    @property
    def num_operons(self):
        return len(self._operons)

    def remove_operons(self):
        self._operons = []

    def operons_to_csv(self, filename):
        _operons.write_rows(filename, _operons.operons_rows(self._operons, self.gene_tables))

    # -- genome.py:335-378
    def infer_regulons(self, threshold=0.5, filename=None):
        results, regulons = _operons.infer_regulons(self._operons, threshold)
        if filename:
            _operons.write_rows(filename, _operons.regulon_rows(results, self.gene_tables))
        return regulons

    def __repr__(self):
        return self._strain_name + ': ' + str([c.accession_number for c in self._chromids])


def get_prior(genome, user_input):
    """cgb/__init__.py:293-350: the user's prior probability of regulation or, without one, the
    expected number of sites G / 2^IC (one per regulated operon) over the number of operons
    predicted with splitting switched off."""
    if user_input.prior_regulation_probability:
        return user_input.prior_regulation_probability
    genome.operon_prediction(1, user_input.operon_prediction_distance_tuning_parameter)
    prior = (genome.length / 2 ** genome.TF_binding_model.IC / genome.num_operons)
    genome.remove_operons()
    return prior
