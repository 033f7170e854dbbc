"""A fixed slice of tools/fuzz_parity.py in the GPU suite (the tool itself runs time-boxed with
fresh seeds): random motifs / ragged sequence sets / slices against the oracle, the gene-level
calls, and the chunked pipeline against the resident path.  Seeds 23 and 56 of the kernel mode are
the degenerate motifs that exposed the two fixed-point accuracy bugs."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(0, 12)) + [23, 56, 214])
def test_kernels_against_oracle(seed):
    import fuzz_parity
    fuzz_parity.one_case(seed)


@pytest.mark.parametrize("block", range(3))
def test_gene_level_calls_against_oracle(block):
    import fuzz_parity
    for seed in range(10 * block, 10 * block + 10):
        fuzz_parity.sites_case(seed)


@pytest.mark.parametrize("block", range(3))
def test_chunked_pipeline_against_resident_path(block):
    import fuzz_parity
    for seed in range(25 * block, 25 * block + 25):
        fuzz_parity.chunks_case(seed)
