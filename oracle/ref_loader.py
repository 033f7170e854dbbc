"""Execute the reference's hot-path files VERBATIM (TEST ORACLE, not product).

Loads ``/root/reference/cgb/{misc,binding_model,pssm_model,site_collection}.py``
unmodified through importlib, under the import shims they need on Python 3.12
(SURVEY.md §8c): ``xrange``, a ``cached_property`` module, a stub ``cgb`` package
(the real ``cgb/__init__.py`` is Python-2-only syntax), and the Biopython stand-in
of oracle/biopython_shim.py.  Nothing is copied into the repo; the reference
tree does not exist on the GPU box, so this module is used only in the build
container, by ``tests/golden/make_golden.py`` (which writes the committed
fixtures) and by the container-only cross-check tests.
"""
from __future__ import annotations

import builtins
import functools
import importlib.util
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CGB_REFERENCE_ROOT", "/root/reference")


def plain_sum(iterable, start=0):
    """Left-to-right ``sum`` as in the reference's Python 2.7 (and every CPython
    before 3.12, whose builtin ``sum`` switched to compensated summation for
    floats).  Injected into the verbatim modules' globals so that
    ``sum(pwm[let][i]*w ...)`` (pssm_model.py:166) and
    ``sum(pssm[l][pos] for l in 'ACGT') / 4`` (pssm_model.py:100) round as they do
    in the reference's environment."""
    total = start
    for x in iterable:
        total = total + x
    return total


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cgb", "pssm_model.py"))


def _install_shims():
    from . import biopython_shim as shim

    if not hasattr(builtins, "xrange"):
        builtins.xrange = range

    if "cached_property" not in sys.modules:
        cp = types.ModuleType("cached_property")
        cp.cached_property = functools.cached_property
        sys.modules["cached_property"] = cp

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "Bio" not in sys.modules:
        bio = mod("Bio")
        bio.__path__ = []
        bio.Seq = mod("Bio.Seq", Seq=shim.Seq)
        jaspar = mod("Bio.motifs.jaspar")
        matrix = mod("Bio.motifs.matrix",
                     PositionWeightMatrix=shim.PositionWeightMatrix,
                     PositionSpecificScoringMatrix=shim.PositionSpecificScoringMatrix,
                     GenericPositionMatrix=shim.GenericPositionMatrix)
        motifs = mod("Bio.motifs", create=shim.create, jaspar=jaspar, matrix=matrix)
        motifs.__path__ = []
        bio.motifs = motifs
        alphabet = mod("Bio.Alphabet", generic_dna="ACGT")
        alphabet.__path__ = []
        alphabet.IUPAC = mod("Bio.Alphabet.IUPAC", unambiguous_dna="ACGT")
        bio.Alphabet = alphabet


def load():
    """Returns a namespace with the verbatim ``PSSMModel``, ``TFBindingModel``
    and ``SiteCollection`` classes of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "cgb_ref" in sys.modules:
        return sys.modules["cgb_ref"]._ns
    _install_shims()
    pkg = types.ModuleType("cgb_ref")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "cgb")]
    sys.modules["cgb_ref"] = pkg

    logger = logging.getLogger("cgb_ref")
    my_logger = types.ModuleType("cgb_ref.my_logger")
    my_logger.my_logger = logger
    sys.modules["cgb_ref.my_logger"] = my_logger

    bio_utils = types.ModuleType("cgb_ref.bio_utils")     # only weblogo is imported
    bio_utils.weblogo = lambda seqs, filename: None
    sys.modules["cgb_ref.bio_utils"] = bio_utils

    def load_verbatim(name):
        path = os.path.join(REFERENCE_ROOT, "cgb", name + ".py")
        spec = importlib.util.spec_from_file_location("cgb_ref." + name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules["cgb_ref." + name] = m
        m.__dict__["sum"] = plain_sum          # Python 2.7 summation order/rounding
        spec.loader.exec_module(m)
        return m

    misc = load_verbatim("misc")
    binding_model = load_verbatim("binding_model")
    pssm_model = load_verbatim("pssm_model")
    site_collection = load_verbatim("site_collection")

    ns = types.SimpleNamespace(
        misc=misc, binding_model=binding_model, pssm_model=pssm_model,
        site_collection=site_collection,
        PSSMModel=pssm_model.PSSMModel,
        TFBindingModel=binding_model.TFBindingModel,
        SiteCollection=site_collection.SiteCollection)
    pkg._ns = ns
    return ns


class _TF(object):
    """Just enough of cgb.protein.Protein for SiteCollection.__init__."""
    accession_number = "TF"


def lexa_collections(ns, test_input_json=None):
    """The six LexA collections of the reference's test_input.json."""
    import json
    path = test_input_json or os.path.join(REFERENCE_ROOT, "test_input.json")
    with open(path) as f:
        d = json.load(f)
    return [ns.SiteCollection(mo["sites"], _TF(), mo["name"]) for mo in d["motifs"]]


# ---------------------------------------------------------------------------
# the genome-side stack: gene.py, operon.py, chromid.py, genome.py, bio_utils.py VERBATIM
# ---------------------------------------------------------------------------
class FakePosition(object):
    """Bio.SeqFeature position: the reference reads ``.position`` (gene.py:25,30)."""

    def __init__(self, position):
        self.position = int(position)


class FeatureLocation(object):
    """Stand-in for Bio.SeqFeature.FeatureLocation (0-based start, exclusive end)."""

    def __init__(self, start, end, strand):
        self.start, self.end, self.strand = FakePosition(start), FakePosition(end), strand


class CompoundLocation(object):
    """Stand-in for Bio.SeqFeature.CompoundLocation: chromid.py:108 skips genes whose
    location is not a plain FeatureLocation."""

    def __init__(self, parts, strand):
        self.parts, self.strand = parts, strand


class FakeFeature(object):
    def __init__(self, type_, location, qualifiers):
        self.type, self.location, self.qualifiers = type_, location, qualifiers

    @property
    def strand(self):
        return self.location.strand


class FakeRecord(object):
    """What chromid.py reads of a Biopython SeqRecord: id, description, seq, features."""

    def __init__(self, accession, description, sequence, features):
        self.id, self.description, self.seq, self.features = accession, description, sequence, features


_RECORDS = {}


def register_record(record):
    """Makes ``Chromid(accession, genome)`` of the verbatim chromid.py find this record
    (entrez_utils.get_genome_record / SeqIO.read are the stubbed network + parser)."""
    _RECORDS[record.id] = record


def load_genome_stack():
    """Namespace of ``load()`` extended with the verbatim Genome / Chromid / Gene / Operon."""
    ns = load()
    if hasattr(ns, "Genome"):
        return ns
    import io

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("cStringIO", StringIO=io.StringIO)
    bio = sys.modules["Bio"]
    bio.SeqIO = mod("Bio.SeqIO", read=lambda handle, fmt: _RECORDS[handle.getvalue()])
    bio.SeqFeature = mod("Bio.SeqFeature", FeatureLocation=FeatureLocation, CompoundLocation=CompoundLocation)
    bio.SeqRecord = mod("Bio.SeqRecord", SeqRecord=FakeRecord)
    blast_mod = mod("Bio.Blast")
    blast_mod.__path__ = []
    blast_mod.NCBIXML = mod("Bio.Blast.NCBIXML")
    mod("cgb_ref.entrez_utils", get_genome_record=lambda acc: acc, get_protein_record=lambda acc: acc)

    def load_verbatim(name):
        path = os.path.join(REFERENCE_ROOT, "cgb", name + ".py")
        spec = importlib.util.spec_from_file_location("cgb_ref." + name, path)
        m = importlib.util.module_from_spec(spec)
        sys.modules["cgb_ref." + name] = m
        m.__dict__["sum"] = plain_sum
        spec.loader.exec_module(m)
        return m

    bio_utils = load_verbatim("bio_utils")          # replaces the weblogo-only stub
    ns.pssm_model.weblogo = bio_utils.weblogo
    blast = load_verbatim("blast")
    protein = load_verbatim("protein")
    gene = load_verbatim("gene")
    operon = load_verbatim("operon")
    chromid = load_verbatim("chromid")
    genome = load_verbatim("genome")
    genome.tqdm = lambda it: it                     # progress bar only
    ns.bio_utils, ns.gene, ns.operon, ns.chromid, ns.genome = bio_utils, gene, operon, chromid, genome
    ns.Genome, ns.Chromid, ns.Gene, ns.Operon = genome.Genome, chromid.Chromid, gene.Gene, operon.Operon
    del blast, protein
    return ns
