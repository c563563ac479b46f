// nm_mgpu.cu -- multi-GPU entries (include/nm_b200_mgpu.h): database-sharded matching with one NCCL all-gather
// of the per-query top-2 records, and frame-sharded batched SIFT.  Host orchestration over the single-GPU C-ABI
// (include/nm_b200.h); the only device code here is what NCCL runs.  SURVEY.md section 8e / 8b.
#include "../../include/nm_b200_mgpu.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

namespace {

inline int nccl_err(ncclResult_t r) { return r == ncclSuccess ? NM_OK : NM_ERR_NCCL_BASE + (int)r; }
#define MG_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return NM_ERR_CUDA_BASE + (int)e_; } while (0)
#define MG_NCCL(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) return NM_ERR_NCCL_BASE + (int)r_; } while (0)

struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

} // namespace

struct nm_mgpu_ctx {
    int world = 0, rank0 = 0;            // ranks in the communicator; rank of local device 0
    int n_local = 0;
    bool own_comms = false;
    std::vector<int> dev;
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> stream;
    // matcher workspace per local device, grown on demand: this rank's records and everybody's
    std::vector<float*> rec, allrec;
    std::vector<size_t> rec_cap;
    int q_groups = 1;                    // query groups Q: world = Q x D, rank r scans query block r / D against shard r % D
    // tracing (local device 0)
    int trace = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // SIFT
    std::vector<nm_sift_ctx*> sift;
    int sift_per_dev = 0, sift_capacity = 0, sift_w = 0, sift_h = 0;
};

extern "C" int nm_mgpu_unique_id(void* id128)
{
    if (!id128) return NM_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) <= NM_MGPU_ID_BYTES, "ncclUniqueId does not fit NM_MGPU_ID_BYTES");
    ncclUniqueId id;
    MG_NCCL(ncclGetUniqueId(&id));
    std::memset(id128, 0, NM_MGPU_ID_BYTES);
    std::memcpy(id128, &id, sizeof(id));
    return NM_OK;
}

static int finish_create(nm_mgpu_ctx* c)
{
    DeviceGuard guard;
    c->stream.assign(c->n_local, nullptr);
    c->rec.assign(c->n_local, nullptr);
    c->allrec.assign(c->n_local, nullptr);
    c->rec_cap.assign(c->n_local, 0);
    for (int d = 0; d < c->n_local; ++d) {
        MG_CUDA(cudaSetDevice(c->dev[d]));
        MG_CUDA(cudaStreamCreateWithFlags(&c->stream[d], cudaStreamNonBlocking));
    }
    MG_CUDA(cudaSetDevice(c->dev[0]));
    for (int i = 0; i < 4; ++i) MG_CUDA(cudaEventCreate(&c->ev[i]));
    return NM_OK;
}

extern "C" int nm_mgpu_create(nm_mgpu_ctx** out, int n_dev, const int* devices, void* const* comms)
{
    if (!out || n_dev <= 0) return NM_ERR_INVALID;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < n_dev) { cudaGetLastError(); return NM_ERR_NO_DEVICE; }
    nm_mgpu_ctx* c = new (std::nothrow) nm_mgpu_ctx();
    if (!c) return NM_ERR_ALLOC;
    c->world = c->n_local = n_dev;
    c->rank0 = 0;
    for (int d = 0; d < n_dev; ++d) c->dev.push_back(devices ? devices[d] : d);
    c->comm.assign(n_dev, nullptr);
    int rc = NM_OK;
    if (comms) {
        for (int d = 0; d < n_dev; ++d) c->comm[d] = static_cast<ncclComm_t>(comms[d]);
    } else {
        c->own_comms = true;
        rc = nccl_err(ncclCommInitAll(c->comm.data(), n_dev, c->dev.data()));
    }
    if (rc == NM_OK) rc = finish_create(c);
    if (rc != NM_OK) { nm_mgpu_destroy(c); return rc; }
    *out = c;
    return NM_OK;
}

extern "C" int nm_mgpu_create_rank(nm_mgpu_ctx** out, int rank, int world, const void* id128, void* comm)
{
    if (!out || world <= 0 || rank < 0 || rank >= world || (!comm && !id128)) return NM_ERR_INVALID;
    nm_mgpu_ctx* c = new (std::nothrow) nm_mgpu_ctx();
    if (!c) return NM_ERR_ALLOC;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); delete c; return NM_ERR_NO_DEVICE; }
    c->world = world; c->rank0 = rank; c->n_local = 1;
    c->dev.push_back(dev);
    c->comm.assign(1, nullptr);
    int rc = NM_OK;
    if (comm) {
        c->comm[0] = static_cast<ncclComm_t>(comm);
    } else {
        c->own_comms = true;
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        rc = nccl_err(ncclCommInitRank(&c->comm[0], world, id, rank));
    }
    if (rc == NM_OK) rc = finish_create(c);
    if (rc != NM_OK) { nm_mgpu_destroy(c); return rc; }
    *out = c;
    return NM_OK;
}

extern "C" int nm_mgpu_destroy(nm_mgpu_ctx* c)
{
    if (!c) return NM_OK;
    DeviceGuard guard;
    for (size_t d = 0; d < c->sift.size(); ++d) {
        if (c->sift[d]) { cudaSetDevice(c->dev[d]); nm_sift_destroy(c->sift[d]); }
    }
    for (int d = 0; d < (int)c->stream.size(); ++d) {
        cudaSetDevice(c->dev[d]);
        if (c->stream[d]) { cudaStreamSynchronize(c->stream[d]); cudaStreamDestroy(c->stream[d]); }
        if (d < (int)c->rec.size() && c->rec[d]) cudaFree(c->rec[d]);
        if (d < (int)c->allrec.size() && c->allrec[d]) cudaFree(c->allrec[d]);
    }
    if (!c->dev.empty()) cudaSetDevice(c->dev[0]);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->own_comms)
        for (ncclComm_t cm : c->comm) if (cm) ncclCommDestroy(cm);
    delete c;
    return NM_OK;
}

extern "C" int nm_mgpu_world(const nm_mgpu_ctx* c) { return c ? c->world : NM_ERR_INVALID; }
extern "C" int nm_mgpu_local(const nm_mgpu_ctx* c) { return c ? c->n_local : NM_ERR_INVALID; }

extern "C" int nm_mgpu_set_trace(nm_mgpu_ctx* c, int enable)
{
    if (!c) return NM_ERR_INVALID;
    c->trace = enable ? 1 : 0;
    return NM_OK;
}

extern "C" int nm_mgpu_match_phase_ms(nm_mgpu_ctx* c, float* ms4)
{
    if (!c || !ms4) return NM_ERR_INVALID;
    DeviceGuard guard;
    MG_CUDA(cudaSetDevice(c->dev[0]));
    MG_CUDA(cudaEventSynchronize(c->ev[3]));
    for (int i = 0; i < 3; ++i) MG_CUDA(cudaEventElapsedTime(&ms4[i], c->ev[i], c->ev[i + 1]));
    MG_CUDA(cudaEventElapsedTime(&ms4[3], c->ev[0], c->ev[3]));
    return NM_OK;
}

extern "C" int nm_mgpu_set_query_groups(nm_mgpu_ctx* c, int q_groups)
{
    if (!c || q_groups <= 0 || c->world % q_groups) return NM_ERR_INVALID;
    c->q_groups = q_groups;
    return NM_OK;
}

extern "C" int nm_mgpu_match_f32(nm_mgpu_ctx* c, const float* const* A_dev, int nA, const float* const* B_dev, const int* nB,
                                 const int* shard_offset, float ambiguity, int* const* match_io_dev, void* const* streams)
{
    if (!c || !A_dev || !B_dev || !nB || !shard_offset || !match_io_dev || nA <= 0) return NM_ERR_INVALID;
    DeviceGuard guard;
    // rank r = (query group r / D, database shard r % D): the rank scans ITS block of the queries against its shard,
    // so the per-row costs of a scan (query packing, seed pass, exact re-rank) shrink with Q as the scan itself does with D
    const int Q = c->q_groups, D = c->world / Q;
    const int nAq = (nA + Q - 1) / Q;                            // rows per query block (the last one may be shorter)
    const size_t blk_floats = (size_t)nAq * 4;
    // workspaces (kept across calls: no allocation on the hot path once they are large enough)
    for (int d = 0; d < c->n_local; ++d) {
        if (!A_dev[d] || !match_io_dev[d] || nB[d] < 0 || (nB[d] > 0 && !B_dev[d])) return NM_ERR_INVALID;
        if (c->rec_cap[d] < blk_floats) {
            MG_CUDA(cudaSetDevice(c->dev[d]));
            if (c->rec[d]) cudaFree(c->rec[d]);
            if (c->allrec[d]) cudaFree(c->allrec[d]);
            c->rec[d] = c->allrec[d] = nullptr; c->rec_cap[d] = 0;
            MG_CUDA(cudaMalloc(&c->rec[d], blk_floats * sizeof(float)));
            MG_CUDA(cudaMalloc(&c->allrec[d], blk_floats * sizeof(float) * c->world));
            c->rec_cap[d] = blk_floats;
        }
    }
    auto st = [&](int d) { return streams ? static_cast<cudaStream_t>(streams[d]) : c->stream[d]; };
    // 1. every device scans its (query block, shard): records (d1, bits(i1 + offset), d2, 0) per query of the block
    for (int d = 0; d < c->n_local; ++d) {
        MG_CUDA(cudaSetDevice(c->dev[d]));
        if (d == 0 && c->trace) MG_CUDA(cudaEventRecord(c->ev[0], st(d)));
        const int q = (c->rank0 + d) / D;
        const int a0 = q * nAq, na = nA - a0 < nAq ? nA - a0 : nAq;
        if (na < nAq) MG_CUDA(cudaMemsetAsync(c->rec[d], 0, blk_floats * sizeof(float), st(d)));   // padding rows of the last block
        if (na > 0) {
            const int rc = nm_match_top2_f32(A_dev[d] + (size_t)a0 * 128, na, B_dev[d], nB[d], shard_offset[d], c->rec[d], st(d));
            if (rc != NM_OK) return rc;
        }
        if (d == 0 && c->trace) MG_CUDA(cudaEventRecord(c->ev[1], st(d)));
    }
    // 2. one all-gather of the 16-byte records: rank-major = (query block, shard)-major, the order the merge needs
    if (c->world > 1) {
        MG_NCCL(ncclGroupStart());
        for (int d = 0; d < c->n_local; ++d) {
            const ncclResult_t r = ncclAllGather(c->rec[d], c->allrec[d], blk_floats, ncclFloat, c->comm[d], st(d));
            if (r != ncclSuccess) { ncclGroupEnd(); return nccl_err(r); }
        }
        MG_NCCL(ncclGroupEnd());
    }
    // 3. per query block: merge its D shard records on every device + the reference's ratio rule (match.cu:88-116)
    for (int d = 0; d < c->n_local; ++d) {
        MG_CUDA(cudaSetDevice(c->dev[d]));
        if (d == 0 && c->trace) MG_CUDA(cudaEventRecord(c->ev[2], st(d)));
        for (int q = 0; q < Q; ++q) {
            const int a0 = q * nAq, na = nA - a0 < nAq ? nA - a0 : nAq;
            if (na <= 0) break;
            const float* recs = c->world > 1 ? c->allrec[d] + (size_t)q * D * blk_floats : c->rec[d];
            // shard s of the block sits nAq rows after shard s - 1 (the last query block may hold fewer than nAq rows)
            const int rc = nm_match_merge_top2_strided(recs, D, nAq, na, ambiguity, match_io_dev[d] + a0, st(d));
            if (rc != NM_OK) return rc;
        }
        if (d == 0 && c->trace) MG_CUDA(cudaEventRecord(c->ev[3], st(d)));
    }
    if (!streams)
        for (int d = 0; d < c->n_local; ++d) {
            MG_CUDA(cudaSetDevice(c->dev[d]));
            MG_CUDA(cudaStreamSynchronize(c->stream[d]));
        }
    return NM_OK;
}

extern "C" int nm_mgpu_sift_create(nm_mgpu_ctx* c, const nm_sift_params* params, int max_frames, int capacity)
{
    if (!c || !params || max_frames <= 0 || capacity <= 0) return NM_ERR_INVALID;
    DeviceGuard guard;
    for (size_t d = 0; d < c->sift.size(); ++d)
        if (c->sift[d]) { cudaSetDevice(c->dev[d]); nm_sift_destroy(c->sift[d]); }
    c->sift.assign(c->n_local, nullptr);
    c->sift_per_dev = (max_frames + c->n_local - 1) / c->n_local;
    c->sift_capacity = capacity; c->sift_w = params->width; c->sift_h = params->height;
    for (int d = 0; d < c->n_local; ++d) {
        MG_CUDA(cudaSetDevice(c->dev[d]));
        const int rc = nm_sift_create(&c->sift[d], params, c->sift_per_dev, capacity);
        if (rc != NM_OK) return rc;
    }
    return NM_OK;
}

extern "C" int nm_mgpu_sift_run_host(nm_mgpu_ctx* c, const float* frames_host, int n_frames, int* counts_host,
                                     float* desc_host, float* x_host, float* y_host)
{
    if (!c || c->sift.empty() || !frames_host || !counts_host || n_frames <= 0 || n_frames > c->sift_per_dev * c->n_local)
        return NM_ERR_INVALID;
    const int q = n_frames / c->n_local, r = n_frames % c->n_local;
    const size_t fpix = (size_t)c->sift_w * c->sift_h, cap = (size_t)c->sift_capacity;
    std::vector<int> rcs(c->n_local, NM_OK);
    std::vector<std::thread> workers;
    for (int d = 0; d < c->n_local; ++d) {
        const int lo = d * q + (d < r ? d : r), n = q + (d < r ? 1 : 0);
        if (n == 0) continue;
        if (n > c->sift_per_dev) return NM_ERR_INVALID;
        workers.emplace_back([=, &rcs] {
            if (cudaSetDevice(c->dev[d]) != cudaSuccess) { rcs[d] = NM_ERR_NO_DEVICE; return; }
            rcs[d] = nm_sift_run_host(c->sift[d], frames_host + lo * fpix, n, counts_host + lo,
                                      desc_host ? desc_host + lo * cap * 128 : nullptr, x_host ? x_host + lo * cap : nullptr,
                                      y_host ? y_host + lo * cap : nullptr, c->stream[d]);
        });
    }
    for (auto& w : workers) w.join();
    for (int rc : rcs) if (rc != NM_OK) return rc;
    return NM_OK;
}
