"""CPU oracle for the cgb hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline.  The product package ``cgb_b200`` never imports it.

Parity status
-------------
* ``pssm_model.py`` / ``binding_model.py`` of the reference are executed
  VERBATIM (``oracle/ref_loader.py``) in the build container to generate the
  golden vectors under ``tests/golden/``; the restatement in
  ``oracle/np_oracle.py`` is checked bit-for-bit against those vectors.
* ``genome.py`` / ``chromid.py`` / ``gene.py`` / ``operon.py`` / ``bio_utils.py`` are executed
  VERBATIM the same way (``ref_loader.load_genome_stack``; network, BLAST and the GenBank
  parser stubbed) on a synthetic annotated genome: gene selection, coordinate rules,
  identified sites + CSV text, posteriors per gene, operons and regulon reports are recorded in
  ``tests/golden/genome_vectors.json`` and pin the oracle's and the product's genome side.
* The arithmetic that lives in Biopython 1.76 (``Bio.motifs``; not vendored in
  ``/root/reference``, not installed, no network) is restated from its
  published algorithm.  The reference ships no tests or golden outputs for
  this path, so that layer is **parity unpinned** (see DESIGN.md).
"""
