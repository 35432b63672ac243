"""CPU oracle for the consensus hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A restatement of the reference's algorithm (shelvick/quoracle, Elixir) for the one path this
repository accelerates.  Every function cites the reference file:line it follows.

Rules of use (enforced by tests/test_layout_rules.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import or execute anything in this package;
  * nothing under ``quoracle_b200/`` imports it -- the product path has no CPU fallback.

Parity status
  * cosine, threshold rule, mode fallback, tie-break, confidence, clustering fingerprints, chunk
    mean, percentile/median/wait rules: PINNED against the reference's own known-answer tests,
    transcribed into ``tests/golden/reference_vectors.json`` by ``tests/golden/make_golden.py``
    (each vector carries the reference test file:line it comes from).
  * embedding VALUES: "parity unpinned".  The reference calls a remote provider
    (lib/quoracle/models/embeddings.ex:277-298, req_llm 1.5.1 in mix.lock); no weights exist in
    the reference.  ``oracle.encoder`` restates THIS build's encoder definition in numpy f64.
  * cl100k token counts (tiktoken 0.4.2, mix.lock): "parity unpinned" -- the vocabulary is not
    available offline; the chunker takes an injectable token counter.
  * The reference cannot be executed in this image (no BEAM), so there is no ``oracle/_ref``.
"""
