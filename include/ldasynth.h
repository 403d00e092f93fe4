/*
 * ldasynth.h -- C ABI of libldasynth.so: the synthetic-corpus generator of the benchmarks and tests.
 * NOT part of the product library (libldagpu.so): bench.py's reference arm and the CPU tests load it
 * without mapping any GPU code.  Corpus shapes: SURVEY.md 8(d); the shapes the reference's authors used are in
 * src/main/resources/datasets/README.txt:3-31 of the reference.
 */
#ifndef LDASYNTH_H
#define LDASYNTH_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* host-side helper for benchmarks and tests: LDA-generative synthetic corpus of a given shape
 * (SURVEY 8d).  doc_offsets int64[D+1] out; tokens int32[capacity] out; returns N in *n_tokens.
 * Tokens of a document are sorted by type id (bag of words, like the bundled corpora).
 * Generates documents [doc_first, doc_first + D) of the corpus that seed defines, so ranks can
 * build their own shard. */
int ldasynth_corpus(int64_t D, int64_t doc_first, int32_t V, int32_t K_gen, double mean_len,
                        double sigma_len, int32_t max_len, uint64_t seed, int64_t *doc_offsets,
                        int32_t *tokens, int64_t capacity, int64_t *n_tokens);

#ifdef __cplusplus
}
#endif
#endif
