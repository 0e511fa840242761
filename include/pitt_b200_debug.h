/*
 * pitt_b200_debug.h — test and measurement hooks of libpitt_b200.so. NOT part of the drop-in boundary
 * (include/pitt_b200.h): nothing a service callback needs is declared here.
 *
 * The kernel-selection hooks below are PROCESS-WIDE switches read at launch time. They exist so that the parity
 * tests can force every variant of a kernel on the same input (tests/test_gpu_plane*.py, test_gpu_primitives.py)
 * and compare it with the oracle; production code never calls them, and they must not be flipped while other
 * contexts are running. The timing hook is per context.
 */
#ifndef PITT_B200_DEBUG_H_
#define PITT_B200_DEBUG_H_

#include <stdint.h>

#include "pitt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* plane scoring variant: 0 automatic (tensor path from 2^27 evaluations), 1 exact packed CUDA-core kernel only,
 * 2 FFMA filter + exact re-evaluation, 3 tensor-core path always */
void pitt_debug_plane_mode(int mode);
/* 1: plane scoring through the generic score_kernel (lanes = points) */
void pitt_debug_force_generic_plane(int on);
/* sphere / cylinder / cone scoring variant: 0 automatic, 1 generic score_kernel, 2 / 3 packed sphere kernel with 512- / 128-point tiles */
void pitt_debug_score_mode(int mode);
/* 1: small selections take the four-launch path instead of select_small_kernel */
void pitt_debug_select_no_fuse(int v);
/* primitive fits of pitt_segment_frame: 0 = batched per model (every launch serves all clusters of a support), 1 = the round-1
 * path (one synchronous seg.segment() per fit), 2 = one asynchronous chain per fit; results are identical
 * (tests/test_gpu_services.py) */
void pitt_debug_frame_mode(int mode);
/* rows from which Levenberg-Marquardt runs on a thread-block cluster */
void pitt_debug_lm_cluster_min(int rows);
/* pitt_sac_segment_host: k equal chunks (1..8), 0 = by cloud size */
void pitt_debug_stream_chunks(int k);
/* statistics of the last scoring call made while collection was enabled: out2[0] = (hypothesis, tile) pairs, out2[1] = pairs
 * re-evaluated exactly */
void pitt_debug_plane_filter_stats(int enable, uint64_t* out2);
void pitt_debug_plane_tc_stats(int enable, uint64_t* out2);
/* per-CTA (SM id << 48 | cycles) + per-phase counters of the last tensor-path call with statistics on (176 entries) */
void pitt_debug_plane_tc_cta_cycles(uint64_t* out176);
/* raw accumulators of hypothesis block 0 x point tile 0 (128 x 256 floats, then sigma and C); returns the float count */
int pitt_debug_plane_tc_dump(int enable, float* out);
/* accumulation error bound of the tensor path in units of u m (default 8; tools/tc_check.py numerics measures 2.62) */
void pitt_debug_plane_tc_acc_ulps(float ulps);
void pitt_debug_plane_tc_variant(int v);
void pitt_debug_plane_tc_nwq(int v);
/* per context: enable = every tensor-path scoring call of ctx records two CUDA events around the plane_tc_kernel launch
 * alone; pitt_debug_plane_tc_kernel_ms waits for the last pair and returns that launch's duration in ms (< 0: none) */
void pitt_debug_plane_tc_time_kernel(pitt_ctx* ctx, int enable);
double pitt_debug_plane_tc_kernel_ms(pitt_ctx* ctx);

/* the device's Philox-4x32-10 (the PHILOX sampler's generator): one raw block for a counter (4 words) and a key (2 words);
 * checked against the Random123 known-answer vectors */
int pitt_debug_philox(pitt_ctx* ctx, const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4);
/* the minimal sample sets the PHILOX sampler draws for hypotheses 0..H-1 of batch stream_id on a cloud of n points (H x S) */
int pitt_debug_philox_samples(pitt_ctx* ctx, int H, int S, int n, uint32_t stream_id, int32_t* out);
/* enable = collect diagnostics in the large-cloud k-NN calls that follow (process wide); out16 (nullable) = those of the last
 * call of ctx: [0] finite points, [1] queries that took the general ring search, [2..5] fast-path attempts at level 0..3,
 * [6] candidates inside the guaranteed radius, [7] / [8] attempts with > 64 candidates below the threshold (first / later
 * bucket), [9] attempts with < k candidates inside the guaranteed radius, [10] most candidates of one attempt */
int pitt_debug_knn_stats(pitt_ctx* ctx, int enable, int64_t* out16);
/* enable != 0: record {start ns, end ns, SM, candidates of the histogram passes, histogram passes, hand-overs} of every CTA of
 * the following knn_collect_kernel launches (whole-cloud normals); out (nullable, 6 * cap words): the record of the last launch. Returns the number of CTAs copied. (knn.cu) */
int pitt_debug_knn_timeline(pitt_ctx* ctx, int enable, uint64_t* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* PITT_B200_DEBUG_H_ */
