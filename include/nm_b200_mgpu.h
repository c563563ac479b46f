/* nm_b200_mgpu.h -- multi-GPU entries of the B200-native NiftyMatch hot path (C-ABI, libnm_b200_mgpu.so).
 *
 * The reference is single-GPU; its public API offers nothing to replace here.  These entries are what a C / C++
 * consumer of the drop-in calls to use the 8 GPUs of a box for the two paths that shard:
 *
 *   - brute-force k=2 matching against a large database (compute_sift_matches, gpu/sift/siftfunctions.cu:15-40):
 *     the DATABASE rows are sharded over the GPUs, the queries replicated; every GPU scans its shard
 *     (nm_match_top2_f32), the 16-byte per-query records (d1, global index, d2) are exchanged with ONE
 *     ncclAllGather over NVLink / NVSwitch and merged on every GPU (nm_match_merge_top2).  The result is
 *     bit-identical to the single-GPU nm_match_f32.
 *   - batched SIFT detect+describe (the client loop of gpu/sift/siftfunctions.h:30-101): FRAMES are split in
 *     contiguous ranges over the GPUs, no collective.
 *
 * Two process models, one context type:
 *   nm_mgpu_create       one process drives n_dev GPUs (communicators from ncclCommInitAll, or the caller's)
 *   nm_mgpu_create_rank  one process per GPU (torchrun / MPI style): rank r of `world` on the CURRENT device
 * Array arguments have one entry per LOCAL device (n_dev, or 1 in the per-rank model).
 *
 * Conventions as in nm_b200.h: 0 = NM_OK, negative = NM_ERR_*, NM_ERR_CUDA_BASE + cudaError_t, and
 * NM_ERR_NCCL_BASE + ncclResult_t for NCCL failures; nothing throws or exits.
 */
#ifndef NM_B200_MGPU_H
#define NM_B200_MGPU_H
#include "nm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NM_ERR_NCCL_BASE 20000
#define NM_MGPU_ID_BYTES 128

typedef struct nm_mgpu_ctx nm_mgpu_ctx;

/* ncclGetUniqueId into id128 (NM_MGPU_ID_BYTES bytes): rank 0 calls it and hands the bytes to the other ranks by its
 * own means (file, socket, MPI_Bcast, torch.distributed broadcast) before they call nm_mgpu_create_rank. */
int nm_mgpu_unique_id(void* id128);

/* One process, n_dev GPUs.  devices: CUDA ordinals (NULL = 0 .. n_dev-1).  comms: n_dev caller-owned ncclComm_t
 * (rank d of n_dev on devices[d]) or NULL, in which case the context creates (ncclCommInitAll) and owns them. */
int nm_mgpu_create(nm_mgpu_ctx** out, int n_dev, const int* devices, void* const* comms);

/* One process per GPU: this process is rank `rank` of `world` and uses the CURRENT device.  comm: the caller's
 * ncclComm_t, or NULL: created from id128 (ncclCommInitRank; collective -- every rank must call it). */
int nm_mgpu_create_rank(nm_mgpu_ctx** out, int rank, int world, const void* id128, void* comm);

int nm_mgpu_destroy(nm_mgpu_ctx* ctx);
int nm_mgpu_world(const nm_mgpu_ctx* ctx);        /* ranks in the communicator */
int nm_mgpu_local(const nm_mgpu_ctx* ctx);        /* devices this process drives */

/* compute_sift_matches against a row-sharded database.  Per local device d:
 *   A_dev[d]         nA x 128 queries on device d (the same values on every device)
 *   B_dev[d], nB[d]  device d's shard: database rows [shard_offset[d], shard_offset[d] + nB[d]) (nB[d] may be 0)
 *   match_io_dev[d]  nA ints on device d, in/out as in the reference (entries with min2 <= 0 stay untouched,
 *                    gpu/kernels/match.cu:107); every device receives the merged result
 *   streams          per-device cudaStream_t to enqueue on, or NULL: the context's own streams, synchronised
 *                    before the call returns.  With caller streams nothing synchronises.
 * Collective: every rank of the communicator must make the call with the same nA. */
int nm_mgpu_match_f32(nm_mgpu_ctx* ctx, const float* const* A_dev, int nA, const float* const* B_dev, const int* nB,
                      const int* shard_offset, float ambiguity, int* const* match_io_dev, void* const* streams);

/* Two-dimensional sharding: the world is cut into q_groups query groups of D = world / q_groups ranks; rank r scans
 * query block r / D (rows [q * ceil(nA / Q), ...) of A) against database shard r % D -- the caller passes, on every
 * rank, the shard of its r % D (the database is cut D ways and held by Q ranks each).  The per-row costs of a scan
 * (query packing, seed pass, exact re-rank of every row) then shrink with Q as the scan itself shrinks with D; the
 * exchange stays ONE all-gather (world blocks of ceil(nA / Q) records) followed by one merge per query block.  q_groups
 * must divide the world; 1 (default) = database sharding only.  The result does not depend on it. */
int nm_mgpu_set_query_groups(nm_mgpu_ctx* ctx, int q_groups);

/* Device milliseconds of the phases of the LAST nm_mgpu_match_f32 on local device 0 when tracing was enabled
 * (nm_mgpu_set_trace(ctx, 1)): {shard scan (pack, tcgen05 scan, exact re-rank, fallback), all-gather, merge, total}. */
int nm_mgpu_set_trace(nm_mgpu_ctx* ctx, int enable);
int nm_mgpu_match_phase_ms(nm_mgpu_ctx* ctx, float* ms4);

/* Batched SIFT over the local devices: one nm_sift context per device for up to ceil(max_frames / n_dev) frames. */
int nm_mgpu_sift_create(nm_mgpu_ctx* ctx, const nm_sift_params* params, int max_frames, int capacity);
/* frames_host: n_frames x height x width floats (pinned memory makes the copies asynchronous); device d takes the
 * d-th contiguous range (the first n_frames % n_dev ranges are one frame longer); outputs as nm_sift_run_host, in
 * frame order: counts[n_frames], desc[n_frames * capacity * 128], x / y[n_frames * capacity] (desc / x / y may be NULL).
 * One host thread per device; returns when all devices are done. */
int nm_mgpu_sift_run_host(nm_mgpu_ctx* ctx, const float* frames_host, int n_frames, int* counts_host,
                          float* desc_host, float* x_host, float* y_host);

#ifdef __cplusplus
}
#endif
#endif /* NM_B200_MGPU_H */
