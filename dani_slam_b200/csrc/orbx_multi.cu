// orbx_multi.cu — the multi-GPU entry points of the C ABI: DB-sharded Hamming kNN with ONE all-gather of
// the per-shard top-2 records over NCCL (NVLink / NVSwitch), and frame batches split over several devices (no collective).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, or the path in ORBX_NCCL_LIB): liborbx.so itself has no link-time
// dependency on it, so single-GPU hosts load the library without NCCL being installed.  Two ways to build communicators:
//   one process per GPU (torchrun, MPI):  rank 0 calls orbx_comm_unique_id, ships the 128 bytes to the other ranks by any
//                                         means, every rank calls orbx_comm_create(world, rank, id, device);
//   one process, several GPUs (the C++ SLAM host): orbx_comm_create_all(comms, n, devices).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "orbx_internal.h"

namespace {

typedef struct { char internal[128]; } NcclUniqueId;     // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)
typedef void *NcclComm;                                    // ncclComm_t
enum { kNcclSuccess = 0, kNcclInt32 = 2 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(NcclComm *, int, const int *) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};

thread_local std::string tl_comm_error;

NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {getenv("ORBX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.err = "NCCL not found (libnccl.so.2; set ORBX_NCCL_LIB to its path)"; return; }
#define ORBX_NCCL_SYM(field, name) \
        *(void **)(&api.field) = dlsym(api.lib, name); \
        if (!api.field) { api.err = std::string("NCCL symbol missing: ") + name; return; }
        ORBX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        ORBX_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        ORBX_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        ORBX_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        ORBX_NCCL_SYM(AllGather, "ncclAllGather")
        ORBX_NCCL_SYM(GroupStart, "ncclGroupStart")
        ORBX_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        ORBX_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef ORBX_NCCL_SYM
    });
    return api.err.empty() ? &api : nullptr;
}

}  // namespace

struct orbx_comm {
    NcclComm comm = nullptr;
    int rank = 0, world = 1, device = 0;
    int32_t *d_rec = nullptr, *d_all = nullptr;   // this rank's packed top-2 record {idx[nq×2], dist[nq×2]} and the gathered records
    int recCap = 0;                               // queries the two buffers hold
    std::string err;
};

namespace {

int comm_fail(orbx_comm *c, const std::string &m) { if (c) c->err = m; else tl_comm_error = m; return ORBX_ERR_CUDA; }

int comm_reserve(orbx_comm *c, int nq) {
    if (nq <= c->recCap && c->d_rec) return ORBX_OK;
    if (c->d_rec) cudaFree(c->d_rec);
    if (c->d_all) cudaFree(c->d_all);
    c->d_rec = c->d_all = nullptr; c->recCap = 0;
    cudaError_t e = cudaMalloc((void **)&c->d_rec, (size_t)nq * 4 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc((void **)&c->d_all, (size_t)c->world * nq * 4 * sizeof(int32_t));
    if (e != cudaSuccess) return comm_fail(c, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    c->recCap = nq;
    return ORBX_OK;
}

}  // namespace

extern "C" {

int orbx_comm_unique_id(uint8_t id128[128]) {
    NcclApi *api = nccl_api();
    if (!api || !id128) { tl_comm_error = api ? "orbx_comm_unique_id: null argument" : nccl_api() ? "" : "NCCL not available"; return ORBX_ERR_ARG; }
    NcclUniqueId id;
    const int rc = api->GetUniqueId(&id);
    if (rc != kNcclSuccess) { tl_comm_error = std::string("ncclGetUniqueId: ") + api->GetErrorString(rc); return ORBX_ERR_CUDA; }
    memcpy(id128, id.internal, 128);
    return ORBX_OK;
}

orbx_comm *orbx_comm_create(int world, int rank, const uint8_t id128[128], int device) {
    NcclApi *api = nccl_api();
    if (!api) { tl_comm_error = "orbx_comm_create: NCCL not available (libnccl.so.2 not found; set ORBX_NCCL_LIB)"; return nullptr; }
    if (world < 1 || rank < 0 || rank >= world || !id128) { tl_comm_error = "orbx_comm_create: bad argument"; return nullptr; }
    OrbxDeviceGuard dg_(device);
    if (dg_.status != cudaSuccess) { tl_comm_error = std::string("cudaSetDevice: ") + cudaGetErrorString(dg_.status); return nullptr; }
    orbx_comm *c = new orbx_comm;
    c->rank = rank; c->world = world; c->device = device;
    NcclUniqueId id;
    memcpy(id.internal, id128, 128);
    const int rc = api->CommInitRank(&c->comm, world, id, rank);
    if (rc != kNcclSuccess) { tl_comm_error = std::string("ncclCommInitRank: ") + api->GetErrorString(rc); delete c; return nullptr; }
    return c;
}

int orbx_comm_create_all(orbx_comm **comms, int n_devices, const int *devices) {
    NcclApi *api = nccl_api();
    if (!api) { tl_comm_error = "orbx_comm_create_all: NCCL not available (libnccl.so.2 not found; set ORBX_NCCL_LIB)"; return ORBX_ERR_CUDA; }
    if (!comms || n_devices < 1 || !devices) { tl_comm_error = "orbx_comm_create_all: bad argument"; return ORBX_ERR_ARG; }
    std::vector<NcclComm> raw(n_devices, nullptr);
    const int rc = api->CommInitAll(raw.data(), n_devices, devices);
    if (rc != kNcclSuccess) { tl_comm_error = std::string("ncclCommInitAll: ") + api->GetErrorString(rc); return ORBX_ERR_CUDA; }
    for (int i = 0; i < n_devices; ++i) {
        comms[i] = new orbx_comm;
        comms[i]->comm = raw[i]; comms[i]->rank = i; comms[i]->world = n_devices; comms[i]->device = devices[i];
    }
    return ORBX_OK;
}

void orbx_comm_destroy(orbx_comm *c) {
    if (!c) return;
    OrbxDeviceGuard dg_(c->device);
    if (c->d_rec) cudaFree(c->d_rec);
    if (c->d_all) cudaFree(c->d_all);
    NcclApi *api = nccl_api();
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
}

const char *orbx_comm_last_error(const orbx_comm *c) { return c ? c->err.c_str() : tl_comm_error.c_str(); }
int orbx_comm_rank(const orbx_comm *c) { return c ? c->rank : -1; }
int orbx_comm_world(const orbx_comm *c) { return c ? c->world : 0; }

// one rank's part: local top-2 → (all-gather) → merge; `phase` lets the single-process form put the gathers of all devices
// into one NCCL group (0: everything, 1: local scan only, 2: all-gather only, 3: merge only)
static int knn2_sharded_phase(orbx_matcher *m, orbx_comm *c, const uint8_t *d_query, int nq, const uint8_t *d_db_shard, int64_t ndb_shard,
                              int64_t idx_base, int32_t *d_idx, int32_t *d_dist, int phase) {
    NcclApi *api = nccl_api();
    if (!api) return comm_fail(c, "NCCL not available");
    if (!m || !c || nq < 0 || ndb_shard < 0 || (nq > 0 && (!d_query || !d_idx || !d_dist))) { if (c) c->err = "orbx_knn2_sharded: bad argument"; return ORBX_ERR_ARG; }
    if (nq == 0) return ORBX_OK;
    OrbxDeviceGuard dg_(c->device);
    if (dg_.status != cudaSuccess) return comm_fail(c, std::string("cudaSetDevice: ") + cudaGetErrorString(dg_.status));
    int rc;
    if (phase == 0 || phase == 1) {
        if ((rc = comm_reserve(c, nq))) return rc;
        rc = orbx_hamming_knn2_device(m, d_query, nq, d_db_shard, ndb_shard, idx_base, c->d_rec, c->d_rec + 2 * (size_t)nq);
        if (rc) return comm_fail(c, std::string("local scan: ") + orbx_matcher_last_error(m)), rc;
    }
    if (phase == 0 || phase == 2) {
        const int nrc = api->AllGather(c->d_rec, c->d_all, (size_t)nq * 4, kNcclInt32, c->comm, (cudaStream_t)orbx_matcher_stream(m));
        if (nrc != kNcclSuccess) return comm_fail(c, std::string("ncclAllGather: ") + api->GetErrorString(nrc));
    }
    if (phase == 0 || phase == 3) {
        rc = orbx_knn2_merge_packed_device(m, c->d_all, c->world, nq, d_idx, d_dist);
        if (rc) return comm_fail(c, std::string("merge: ") + orbx_matcher_last_error(m)), rc;
    }
    return ORBX_OK;
}

int orbx_knn2_sharded(orbx_matcher *m, orbx_comm *c, const uint8_t *d_query, int nq, const uint8_t *d_db_shard, int64_t ndb_shard, int64_t idx_base,
                      int32_t *d_idx, int32_t *d_dist) {
    return knn2_sharded_phase(m, c, d_query, nq, d_db_shard, ndb_shard, idx_base, d_idx, d_dist, 0);
}

int orbx_knn2_sharded_all(orbx_matcher *const *ms, orbx_comm *const *cs, int n, const uint8_t *const *d_query, int nq, const uint8_t *const *d_db_shard,
                          const int64_t *ndb_shard, const int64_t *idx_base, int32_t *const *d_idx, int32_t *const *d_dist) {
    NcclApi *api = nccl_api();
    if (!api || !ms || !cs || n < 1 || !d_query || !d_db_shard || !ndb_shard || !idx_base || !d_idx || !d_dist) { tl_comm_error = "orbx_knn2_sharded_all: bad argument or NCCL not available"; return ORBX_ERR_ARG; }
    int rc = ORBX_OK;
    for (int i = 0; i < n && !rc; ++i) rc = knn2_sharded_phase(ms[i], cs[i], d_query[i], nq, d_db_shard[i], ndb_shard[i], idx_base[i], d_idx[i], d_dist[i], 1);
    if (rc) return rc;
    api->GroupStart();                                   // one thread drives several devices: the collective calls must be grouped
    for (int i = 0; i < n && !rc; ++i) rc = knn2_sharded_phase(ms[i], cs[i], d_query[i], nq, d_db_shard[i], ndb_shard[i], idx_base[i], d_idx[i], d_dist[i], 2);
    const int grc = api->GroupEnd();
    if (rc) return rc;
    if (grc != kNcclSuccess) { tl_comm_error = std::string("ncclGroupEnd: ") + api->GetErrorString(grc); return ORBX_ERR_CUDA; }
    for (int i = 0; i < n && !rc; ++i) rc = knn2_sharded_phase(ms[i], cs[i], d_query[i], nq, d_db_shard[i], ndb_shard[i], idx_base[i], d_idx[i], d_dist[i], 3);
    return rc;
}

// Frame batches over several devices: contiguous slices of the batch, one host thread and one extractor handle per device, no
// collective.  Arguments are those of orbx_extract_batch; handle i gets frames [i·batch/n, (i+1)·batch/n).
int orbx_extract_batch_multi(orbx_extractor *const *exs, int n_handles, const uint8_t *const *images, int batch, int rows, int cols, size_t step,
                             const int32_t *rects_xywh, int n_rects, int lap0, int lap1, orbx_keypoint *keypoints, uint8_t *descriptors, int cap,
                             int32_t *n_out, int32_t *mono_index) {
    if (!exs || n_handles < 1 || batch < 0) return ORBX_ERR_ARG;
    for (int i = 0; i < n_handles; ++i) if (!exs[i]) return ORBX_ERR_ARG;
    std::vector<int> rcs(n_handles, ORBX_OK);
    std::vector<std::thread> th;
    auto slice = [&](int i) {
        const int lo = (int)((long long)batch * i / n_handles), hi = (int)((long long)batch * (i + 1) / n_handles);
        if (hi > lo)
            rcs[i] = orbx_extract_batch(exs[i], images + lo, hi - lo, rows, cols, step, rects_xywh, n_rects, lap0, lap1, keypoints + (size_t)lo * cap,
                                        descriptors + (size_t)lo * cap * 32, cap, n_out + lo, mono_index + lo);
    };
    for (int i = 1; i < n_handles; ++i) th.emplace_back(slice, i);
    slice(0);
    for (auto &t : th) t.join();
    int rc = ORBX_OK;
    for (int i = 0; i < n_handles; ++i)                          // the first hard error wins; a capacity overflow is reported if nothing worse happened
        if (rcs[i] != ORBX_OK && (rc == ORBX_OK || rc == ORBX_ERR_CAPACITY)) rc = rcs[i];
    return rc;
}

}  // extern "C"
