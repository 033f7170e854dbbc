"""TFBindingModel: abstract TF-binding model + Bayesian regulation posterior.

Drop-in for the reference's ``cgb/binding_model.py:12-138``: same constructor,
properties, ``build_bayesian_estimator`` and ``binding_probability`` contract.
The per-sequence mixture likelihood ratio (binding_model.py:94-113) is not
evaluated here with scipy: subclasses run it as a fused CUDA reduction
(``_posterior_of_sequences``); there is no CPU fallback.
"""
from __future__ import annotations

import logging
from abc import ABCMeta, abstractmethod

import numpy as np

my_logger = logging.getLogger("cgb_b200")


class TFBindingModel(metaclass=ABCMeta):
    def __init__(self, collections,
                 background={'A': 0.25, 'C': 0.25, 'G': 0.25, 'T': 0.25}, alphabet="ACGT"):
        """binding_model.py:35-44."""
        self._background = background
        self._collection_set = collections
        self._alphabet = alphabet

    @property
    def alphabet(self):
        """binding_model.py:46-49."""
        return self._alphabet

    @property
    def background(self):
        """binding_model.py:51-54."""
        return self._background

    @property
    def site_collections(self):
        """binding_model.py:56-59."""
        return self._collection_set

    @property
    def bayesian_estimator(self):
        """binding_model.py:61-69: reads an attribute the reference never assigns,
        so it raises AttributeError there too."""
        return self._bayesian_estimator

    def build_bayesian_estimator(self, bg_scores):
        """binding_model.py:71-92: N(mu_m, std_m) from the model's own sites,
        N(mu_bg, std_bg) from ``bg_scores`` (np.mean / np.std, ddof=0)."""
        my_logger.debug("Building Bayesian estimator.")
        pssm_scores = self.score_self()
        self._mu_m, self._std_m = np.mean(pssm_scores), np.std(pssm_scores)
        self._mu_bg, self._std_bg = np.mean(bg_scores), np.std(bg_scores)
        self._estimator_changed()
        my_logger.info("Regulation parameters: %.2f %.2f" % (self._mu_m, self._std_m))
        my_logger.info("Background parameters: %.2f %.2f" % (self._mu_bg, self._std_bg))

    def binding_probability(self, seq, p_motif, alpha=1/350.0):
        """binding_model.py:94-113 for one sequence (one fused kernel launch)."""
        return float(self._posterior_of_sequences([seq], p_motif, alpha)[0])

    def _estimator_changed(self):
        pass

    @abstractmethod
    def _posterior_of_sequences(self, seqs, p_motif, alpha):
        """Posterior of regulation for each ASCII sequence (device path)."""

    @abstractmethod
    def threshold(self):
        """binding_model.py:116-122."""

    @abstractmethod
    def score_seq(self):
        """binding_model.py:124-130."""

    @property
    @abstractmethod
    def length(self):
        """binding_model.py:132-138."""
