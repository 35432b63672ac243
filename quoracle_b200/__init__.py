"""quoracle_b200 -- B200-native consensus hot path for shelvick/quoracle.

Only what the path needs: the CUDA kernels + C ABI (``csrc/``, ``include/qconsensus.h``), the ctypes
binding (``native``), synthetic inputs (``synth``) and the host-side mirror of the reference's
seams (``embedding_fn``, ``consensus_rules``, ``collector``, ``sharding``).  No CPU fallback.
"""
__version__ = "0.1.0"
