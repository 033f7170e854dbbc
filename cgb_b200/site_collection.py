"""SiteCollection: binding sites of one TF in one species -> PSWM.

Drop-in for the reference's ``cgb/site_collection.py:7-64`` on the hot path (the
JASPAR writer is out of scope).  No Biopython: counts, pseudocounts and column
normalisation come from :mod:`cgb_b200.motif_matrix`.
"""
from __future__ import annotations

from .motif_matrix import counts_from_instances, normalize_counts


class SiteCollection:
    def __init__(self, sites, TF=None, name="", pseudocounts=1):
        """site_collection.py:15-21 (``TF`` is only used for a display name)."""
        self._TF = TF
        self._name = name
        self._instances = [str(site) for site in sites]
        self._counts = counts_from_instances(self._instances)
        self._pseudocounts = pseudocounts
        self._length = len(self._instances[0])
        # the sites back to back as bytes (what the host scorer takes), built once: collections are
        # shared by the many models of a run (one per genome in the reference's go() loop)
        self._uniform = all(len(s) == self._length for s in self._instances)
        self._raw = "".join(self._instances).encode("ascii", "replace") if self._uniform else None

    @property
    def TF(self):
        return self._TF

    @property
    def name(self):
        return self._name

    @property
    def pwm(self):
        """site_collection.py:33-36.  (Computed once: the matrix is immutable -- tuples per letter --
        and every model built on the collection asks for it again.)"""
        pwm = self.__dict__.get("_pwm_cache")
        if pwm is None:
            pwm = self.__dict__["_pwm_cache"] = normalize_counts(self._counts, self._pseudocounts)
        return pwm

    @property
    def IC(self):
        """site_collection.py:38-41."""
        return self.pwm.log_odds(None).mean()

    @property
    def sites(self):
        """site_collection.py:43-46."""
        return list(self._instances)

    @property
    def site_count(self):
        return len(self._instances)

    @property
    def length(self):
        return self._length
