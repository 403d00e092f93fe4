"""ldagroupedgibbssampler_b200 -- B200-native Gibbs-sweep engine (libldagpu.so) behind the sampler
interface of clintpgeorge/LDAGroupedGibbsSampler, for the new schemes ``gpu_ggs`` and ``gpu_pcgs`` (and the sparse ``gpu_spalias`` / ``gpu_polyaurn``).

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/ldagpu.h) and the host-side
mirror of the reference's sampler interface.  There is no CPU fallback: importing works anywhere,
but every compute call needs the built library and a CUDA device.
"""
from ._lib import LdaGpuError, SO_PATH, SYMBOLS, load, synth_corpus  # noqa: F401
from .corpus import (Alphabet, InstanceList, SHAPES, corpus_statistics, load_dataset,  # noqa: F401
                     shard_documents_by_tokens, take_shard, tfidf_ranking, tokenize, write_synthetic_corpus)
from .sampler import GpuLDASampler, LDAConfiguration, SCHEMES, createModel  # noqa: F401

__all__ = ["GpuLDASampler", "LDAConfiguration", "createModel", "InstanceList", "Alphabet", "load_dataset",
           "tokenize", "corpus_statistics", "write_synthetic_corpus", "tfidf_ranking", "synth_corpus", "shard_documents_by_tokens", "take_shard", "SHAPES", "SCHEMES", "LdaGpuError", "load"]
