/*
 * kombgpu_debug.h -- measurement aids of libkombgpu.so.  NOT part of the product ABI (include/kombgpu.h):
 * nothing here replaces anything in the reference; the probes under tools/ and one sort test use them.
 */
#ifndef KOMBGPU_DEBUG_H
#define KOMBGPU_DEBUG_H

#include "kombgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* tools/sort_probe.py: sorts n synthetic 64-bit keys with lo_bits random low bits and hi_bits bits at
 * position 32 (random, or non-decreasing when sorted_hi != 0: the shape of hit arrays in read order) on
 * bits [0, lo_bits) + [32, 32 + hi_bits), `reps` times; returns the best time of the sort alone and
 * whether the result is a sorted permutation of the input. */
int kombgpu_debug_sort_u64(kombgpu_ctx *ctx, uint64_t n, int lo_bits, int hi_bits, int sorted_hi, int reps,
                           float *ms_best, int *ok);

/* tools/peel_ab.py, tools/peel_trace*.py: run the k-core peel again on a graph that already holds its
 * coreness (kombgpu_coreness peels once per graph), so one build serves many timed peels. */
int kombgpu_debug_peel_again(kombgpu_graph *g);

#ifdef __cplusplus
}
#endif
#endif /* KOMBGPU_DEBUG_H */
