"""cgb_b200 — B200-native PSWM scanning and regulation posteriors (the hot path of
ErillLab/cgb): ``pssm_model.PSSMModel`` / ``binding_model.TFBindingModel`` call
surface over hand-written sm_100a kernels behind a C ABI (include/cgbk.h).

Importing the package does not need a GPU; any scoring call does (no CPU fallback).
"""
from .binding_model import TFBindingModel
from .gene_regions import GeneTable, python_slice_bounds
from .genome_buffer import HostPackedGenome, PackedGenome
from .pssm_model import PSSMModel, prepare_thresholds, scan_hits_many
from .site_collection import SiteCollection
from .site_search import Site, calculate_regulation_probabilities, identify_sites
from .genome import Chromid, Genome, get_prior
from . import genbank, operons, motif_batch
from .motif_batch import ModelBatch, scan_batch, unpack_hits

__all__ = ["TFBindingModel", "PSSMModel", "SiteCollection", "PackedGenome", "HostPackedGenome", "GeneTable",
           "Site", "Genome", "ModelBatch", "scan_batch", "unpack_hits", "motif_batch", "Chromid", "get_prior", "prepare_thresholds", "genbank", "operons", "identify_sites", "scan_hits_many", "calculate_regulation_probabilities", "python_slice_bounds"]
