"""voitta-rag_b200 — B200-native (sm_100a) backend for voitta-rag's retrieval hot path.

  vector_store  drop-in for voitta.services.vector_store (VectorStoreService, ChunkMetadata, ...)
  engine        ctypes binding of libvoitta_b200.so (C ABI in include/voitta_b200.h)
  sharded       row-sharded multi-GPU search (one process per GPU, NCCL all-gather of candidates)
  synth         seeded synthetic corpora of BASELINE.json's shapes (bench / tests)
  csrc/         the CUDA kernels and the C ABI

Import as ``voitta_rag_b200`` (shim package next to this directory).
"""
__version__ = "0.1.0"
