// Multi-GPU exchanges of the path, inside the C ABI: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The path shards without a data-path collective: images are independent and label rows
// combine by integer sums.  Only two small exchanges exist, both on the caller's stream:
//   * dedupe across ranks  — all-gather of the 32-byte digests (+ global listing index + validity) and the SAME
//     deterministic resolution on every rank: first occurrence = smallest global listing index, i.e. the
//     reference's sequential "first seen wins" (app/services/webdav_sync.py:324-354) whatever the sharding;
//   * label aggregation    — all-reduce (sum, int64) of the k class totals + integer Fleiss partials; kappa is
//     computed from the integers afterwards, so it is bit-identical for any GPU count.
// NCCL is bound at run time (dlopen) so that single-GPU users of libb2ingest.so do not need it: the copy already
// loaded in the process is preferred (PyTorch ships its own libnccl.so.2), then the system library; B2_NCCL_LIB
// names another one.  The 128-byte unique id travels between the processes by whatever the host has (the
// reference-side binding in INTEGRATION.md uses a file; dist.py uses the torch.distributed store).
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace b2 {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;   // optional (NCCL >= 2.14)
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    char error[256] = "";
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("B2_NCCL_LIB");
        void *h = nullptr;
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);     // the copy this process already uses
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!h) {
            snprintf(api.error, sizeof(api.error), "cannot load libnccl.so.2 (%s); set B2_NCCL_LIB", dlerror());
            return;
        }
        api.handle = h;
        bool ok = true;
#define B2_SYM(field, name)                                                            \
        api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name));             \
        ok = ok && api.field != nullptr
        B2_SYM(GetUniqueId, "ncclGetUniqueId");
        B2_SYM(CommInitRank, "ncclCommInitRank");
        B2_SYM(CommDestroy, "ncclCommDestroy");
        B2_SYM(AllGather, "ncclAllGather");
        B2_SYM(AllReduce, "ncclAllReduce");
        B2_SYM(GroupStart, "ncclGroupStart");
        B2_SYM(GroupEnd, "ncclGroupEnd");
        B2_SYM(GetErrorString, "ncclGetErrorString");
        B2_SYM(GetVersion, "ncclGetVersion");
#undef B2_SYM
        api.CommInitRankConfig = reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
        if (!ok) {
            snprintf(api.error, sizeof(api.error), "libnccl.so.2 lacks a required symbol");
            api.handle = nullptr;
        }
    });
    return api.handle ? &api : nullptr;
}

#define B2_NCCL_CHECK(expr)                                                              \
    do {                                                                                 \
        ncclResult_t _r = (expr);                                                        \
        if (_r != ncclSuccess)                                                           \
            return ::b2::fail(B2_ERR_NCCL, "%s failed: %s (%s:%d)", #expr,               \
                              api->GetErrorString(_r), __FILE__, __LINE__);              \
    } while (0)

// first_seq[i] = global listing index of the first / last occurrence of local image i's content
__global__ void __launch_bounds__(256)
global_slice_kernel(const uint8_t *__restrict__ is_new_all, const int32_t *__restrict__ first_all,
                    const int32_t *__restrict__ last_all, const uint32_t *__restrict__ seq_all, uint32_t base,
                    uint32_t n_local, uint8_t *__restrict__ is_new, int64_t *__restrict__ first_seq,
                    int64_t *__restrict__ last_seq) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    is_new[i] = is_new_all[base + i];
    const int32_t f = first_all[base + i], l = last_all[base + i];
    if (first_seq) first_seq[i] = f < 0 ? -1 : int64_t(seq_all[f]);
    if (last_seq) last_seq[i] = l < 0 ? -1 : int64_t(seq_all[l]);
}

struct GlobalLayout {
    uint64_t send_dig, send_seq, send_val, all_dig, all_seq, all_val, is_new, first, last, ws, ws_bytes, total;
};
static GlobalLayout global_layout(uint32_t world, uint32_t n_max) {
    auto up = [](uint64_t x) { return (x + 255) & ~uint64_t(255); };
    const uint64_t n_all = uint64_t(world) * n_max;
    GlobalLayout g;
    uint64_t o = 0;
    g.send_dig = o; o += up(uint64_t(n_max) * 32);
    g.send_seq = o; o += up(uint64_t(n_max) * 4);
    g.send_val = o; o += up(n_max);
    g.all_dig = o; o += up(n_all * 32);
    g.all_seq = o; o += up(n_all * 4);
    g.all_val = o; o += up(n_all);
    g.is_new = o; o += up(n_all);
    g.first = o; o += up(n_all * 4);
    g.last = o; o += up(n_all * 4);
    g.ws = o;
    g.ws_bytes = b2_dedupe_workspace_bytes(uint32_t(n_all));
    g.total = o + up(g.ws_bytes);
    return g;
}

}  // namespace b2

struct b2_comm {
    ncclComm_t comm = nullptr;
    int device = 0, rank = 0, world = 1;
    // peer reduce (b2_comm_enable_peer_reduce): my mailbox, the peers' mapped through CUDA IPC, the device descriptor
    b2::PeerReduceDesc *d_peer = nullptr;
    uint8_t *d_mailbox = nullptr;
    void *peer_maps[b2::kPeerMaxWorld] = {};
    uint32_t peer_cap = 0;
};

extern "C" int b2_comm_unique_id(uint8_t *id_out) {
    using namespace b2;
    B2_REQUIRE(id_out != nullptr, "b2_comm_unique_id: null output");
    static_assert(sizeof(ncclUniqueId) == B2_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    NcclApi *api = nccl_api();
    if (!api) return fail(B2_ERR_NCCL, "b2_comm_unique_id: %s", nccl_api() ? "" : "NCCL is not available in this process");
    ncclUniqueId id;
    B2_NCCL_CHECK(api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return B2_OK;
}

extern "C" int b2_comm_init(int device, int rank, int world, const uint8_t *id, b2_comm **out) {
    using namespace b2;
    B2_REQUIRE(out != nullptr, "b2_comm_init: null output");
    *out = nullptr;
    B2_REQUIRE(id != nullptr && world >= 1 && rank >= 0 && rank < world, "b2_comm_init: bad rank / world / id");
    int rc = b2_init(device);
    if (rc != B2_OK) return rc;
    NcclApi *api = nccl_api();
    if (!api) return fail(B2_ERR_NCCL, "b2_comm_init: NCCL is not available in this process (libnccl.so.2 not found; set B2_NCCL_LIB)");
    b2_comm *c = new (std::nothrow) b2_comm();
    B2_REQUIRE(c != nullptr, "b2_comm_init: out of host memory");
    c->device = device; c->rank = rank; c->world = world;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    // The exchanges of this path are tiny (8.6 KB of partials, 32 bytes per image of digests): a communicator of at
    // most two CTAs keeps NCCL's kernel off the SMs the tally / hash kernels running beside it need (the default claims
    // up to 32 CTAs and pushed the overlapped tally into a second wave).  B2_NCCL_MAX_CTAS overrides; 0 = NCCL's default.
    int max_ctas = 2;
    if (const char *e = getenv("B2_NCCL_MAX_CTAS")) max_ctas = atoi(e);
    ncclResult_t r;
    if (api->CommInitRankConfig && max_ctas > 0) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = max_ctas;
        r = api->CommInitRankConfig(&c->comm, world, uid, rank, &cfg);
    } else {
        r = api->CommInitRank(&c->comm, world, uid, rank);
    }
    if (r != ncclSuccess) {
        delete c;
        return fail(B2_ERR_NCCL, "ncclCommInitRank failed: %s", api->GetErrorString(r));
    }
    *out = c;
    return B2_OK;
}

extern "C" int b2_comm_destroy(b2_comm *c) {
    using namespace b2;
    if (!c) return B2_OK;
    NcclApi *api = nccl_api();
    cudaSetDevice(c->device);
    if (c->d_peer) {
        cudaDeviceSynchronize();
        for (int q = 0; q < c->world; ++q)
            if (c->peer_maps[q]) cudaIpcCloseMemHandle(c->peer_maps[q]);
        cudaFree(c->d_peer);
        cudaFree(c->d_mailbox);
    }
    if (api && c->comm) api->CommDestroy(c->comm);
    delete c;
    return B2_OK;
}

extern "C" int b2_comm_info(const b2_comm *c, int *rank, int *world, int *nccl_version) {
    using namespace b2;
    B2_REQUIRE(c != nullptr, "b2_comm_info: null communicator");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) {
        NcclApi *api = nccl_api();
        *nccl_version = 0;
        if (api) api->GetVersion(nccl_version);
    }
    return B2_OK;
}

extern "C" int b2_allgather_digests(b2_comm *c, const uint8_t *d_local, uint32_t n_per_rank, uint8_t *d_all, void *stream) {
    using namespace b2;
    B2_REQUIRE(c && (n_per_rank == 0 || (d_local && d_all)), "b2_allgather_digests: null pointer");
    if (n_per_rank == 0) return B2_OK;
    NcclApi *api = nccl_api();
    if (!api) return fail(B2_ERR_NCCL, "b2_allgather_digests: NCCL is not available");
    B2_NCCL_CHECK(api->AllGather(d_local, d_all, size_t(n_per_rank) * 32, ncclUint8, c->comm, static_cast<cudaStream_t>(stream)));
    return B2_OK;
}

extern "C" int b2_allreduce_i64(b2_comm *c, int64_t *d_values, uint64_t count, void *stream) {
    using namespace b2;
    B2_REQUIRE(c && (count == 0 || d_values), "b2_allreduce_i64: null pointer");
    if (count == 0) return B2_OK;
    NcclApi *api = nccl_api();
    if (!api) return fail(B2_ERR_NCCL, "b2_allreduce_i64: NCCL is not available");
    B2_NCCL_CHECK(api->AllReduce(d_values, d_values, size_t(count), ncclInt64, ncclSum, c->comm, static_cast<cudaStream_t>(stream)));
    return B2_OK;
}

extern "C" uint64_t b2_dedupe_global_workspace_bytes(uint32_t world, uint32_t n_max) {
    return b2::global_layout(world ? world : 1, n_max ? n_max : 1).total;
}

extern "C" int b2_dedupe_global(b2_comm *c, const uint8_t *d_digests, const uint8_t *d_valid, const uint32_t *d_seq,
                                uint32_t n_local, uint32_t n_max, const uint8_t *d_existing, uint64_t m,
                                uint8_t *d_is_new, int64_t *d_first_seq, int64_t *d_last_seq, uint32_t *d_counts,
                                void *d_workspace, uint64_t workspace_bytes, void *stream) {
    using namespace b2;
    B2_REQUIRE(c != nullptr && d_counts != nullptr && d_workspace != nullptr, "b2_dedupe_global: null pointer");
    B2_REQUIRE(n_local <= n_max && n_max >= 1, "b2_dedupe_global: n_local must not exceed n_max (the largest shard of any rank)");
    B2_REQUIRE(n_local == 0 || (d_digests && d_seq && d_is_new), "b2_dedupe_global: null pointer");
    B2_REQUIRE(uint64_t(c->world) * n_max < 0x7fffffffull, "b2_dedupe_global: too many digests");
    B2_REQUIRE((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "b2_dedupe_global: workspace must be 256-byte aligned");
    const GlobalLayout g = global_layout(uint32_t(c->world), n_max);
    if (workspace_bytes < g.total)
        return fail(B2_ERR_WORKSPACE, "b2_dedupe_global: workspace %llu < required %llu bytes",
                    (unsigned long long)workspace_bytes, (unsigned long long)g.total);
    NcclApi *api = nccl_api();
    if (!api) return fail(B2_ERR_NCCL, "b2_dedupe_global: NCCL is not available");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t *w = static_cast<uint8_t *>(d_workspace);
    // this rank's shard padded to n_max entries; pad entries are invalid (not counted, never a first occurrence)
    B2_CUDA_CHECK(cudaMemsetAsync(w + g.send_dig, 0, size_t(n_max) * 32, st));
    B2_CUDA_CHECK(cudaMemsetAsync(w + g.send_seq, 0xff, size_t(n_max) * 4, st));
    B2_CUDA_CHECK(cudaMemsetAsync(w + g.send_val, 0, n_max, st));
    if (n_local) {
        B2_CUDA_CHECK(cudaMemcpyAsync(w + g.send_dig, d_digests, size_t(n_local) * 32, cudaMemcpyDeviceToDevice, st));
        B2_CUDA_CHECK(cudaMemcpyAsync(w + g.send_seq, d_seq, size_t(n_local) * 4, cudaMemcpyDeviceToDevice, st));
        if (d_valid) B2_CUDA_CHECK(cudaMemcpyAsync(w + g.send_val, d_valid, n_local, cudaMemcpyDeviceToDevice, st));
        else B2_CUDA_CHECK(cudaMemsetAsync(w + g.send_val, 1, n_local, st));
    }
    B2_NCCL_CHECK(api->GroupStart());
    ncclResult_t r1 = api->AllGather(w + g.send_dig, w + g.all_dig, size_t(n_max) * 32, ncclUint8, c->comm, st);
    ncclResult_t r2 = api->AllGather(w + g.send_seq, w + g.all_seq, size_t(n_max), ncclUint32, c->comm, st);
    ncclResult_t r3 = api->AllGather(w + g.send_val, w + g.all_val, size_t(n_max), ncclUint8, c->comm, st);
    ncclResult_t r4 = api->GroupEnd();
    for (ncclResult_t r : {r1, r2, r3, r4})
        if (r != ncclSuccess) return fail(B2_ERR_NCCL, "b2_dedupe_global: all-gather failed: %s", api->GetErrorString(r));
    const uint32_t n_all = uint32_t(c->world) * n_max;
    int rc = b2_dedupe(w + g.all_dig, w + g.all_val, reinterpret_cast<const uint32_t *>(w + g.all_seq), n_all, d_existing, m,
                       w + g.is_new, reinterpret_cast<int32_t *>(w + g.first), reinterpret_cast<int32_t *>(w + g.last),
                       d_counts, w + g.ws, g.ws_bytes, st);
    if (rc != B2_OK) return rc;
    if (n_local) {
        global_slice_kernel<<<(n_local + 255) / 256, 256, 0, st>>>(
            w + g.is_new, reinterpret_cast<const int32_t *>(w + g.first), reinterpret_cast<const int32_t *>(w + g.last),
            reinterpret_cast<const uint32_t *>(w + g.all_seq), uint32_t(c->rank) * n_max, n_local, d_is_new, d_first_seq, d_last_seq);
        B2_LAUNCH_CHECK("global_slice_kernel");
    }
    return B2_OK;
}

// ---- all-reduce over NVLink peer memory, fused into the tally ------------------------------------------------------
namespace b2 {
int label_tally_impl(const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                     uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                     int32_t *d_counts, int64_t *d_partials, int64_t *d_agree_hist, PeerReduceDesc *peer, void *stream);

__global__ void __launch_bounds__(256) peer_allreduce_kernel(PeerReduceDesc *d, unsigned long long *vec, uint32_t n) {
    peer_allreduce_cta(d, vec, n);
}
}  // namespace b2

// Collective and blocking (every rank calls it once, with the same max_values): allocates this rank's mailbox, exchanges
// the CUDA IPC handles through the communicator and maps the peers' mailboxes.  One process per GPU, all GPUs of one box
// with peer access (NVLink / NVSwitch).
namespace b2 {
static int enable_peer_reduce_impl(b2_comm *c, uint32_t max_values);
}
extern "C" int b2_comm_enable_peer_reduce(b2_comm *c, uint32_t max_values) {
    const int rc = b2::enable_peer_reduce_impl(c, max_values);
    if (rc != B2_OK && c && c->d_peer == nullptr) {          // failed half way: give back what was allocated and mapped
        for (int q = 0; q < c->world && q < b2::kPeerMaxWorld; ++q)
            if (c->peer_maps[q]) { cudaIpcCloseMemHandle(c->peer_maps[q]); c->peer_maps[q] = nullptr; }
        cudaFree(c->d_mailbox);
        c->d_mailbox = nullptr;
        c->peer_cap = 0;
        cudaGetLastError();
    }
    return rc;
}
static int b2::enable_peer_reduce_impl(b2_comm *c, uint32_t max_values) {
    using namespace b2;
    B2_REQUIRE(c != nullptr && max_values >= 1 && max_values <= (1u << 20), "b2_comm_enable_peer_reduce: bad argument");
    B2_REQUIRE(c->world <= kPeerMaxWorld, "b2_comm_enable_peer_reduce: at most %d ranks", kPeerMaxWorld);
    if (c->d_peer) return c->peer_cap >= max_values ? B2_OK : fail(B2_ERR_BAD_ARG, "b2_comm_enable_peer_reduce: already enabled with a smaller capacity");
    NcclApi *api = nccl_api();
    if (!api) return fail(B2_ERR_NCCL, "b2_comm_enable_peer_reduce: NCCL is not available");
    B2_CUDA_CHECK(cudaSetDevice(c->device));
    const uint32_t cap = (max_values + 31u) & ~31u;
    const size_t flag_bytes = 256, mail_bytes = size_t(2) * c->world * cap * 8;   // 2 x world flags fit 256 bytes (world <= 16)
    B2_CUDA_CHECK(cudaMalloc(&c->d_mailbox, flag_bytes + mail_bytes));
    B2_CUDA_CHECK(cudaMemset(c->d_mailbox, 0, flag_bytes + mail_bytes));
    cudaIpcMemHandle_t mine;
    B2_CUDA_CHECK(cudaIpcGetMemHandle(&mine, c->d_mailbox));
    uint8_t *d_handles = nullptr;
    B2_CUDA_CHECK(cudaMalloc(&d_handles, size_t(c->world + 1) * sizeof(mine)));
    B2_CUDA_CHECK(cudaMemcpy(d_handles + size_t(c->world) * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice));
    B2_NCCL_CHECK(api->AllGather(d_handles + size_t(c->world) * sizeof(mine), d_handles, sizeof(mine), ncclUint8, c->comm, nullptr));
    B2_CUDA_CHECK(cudaStreamSynchronize(nullptr));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    B2_CUDA_CHECK(cudaMemcpy(all.data(), d_handles, size_t(c->world) * sizeof(mine), cudaMemcpyDeviceToHost));
    B2_CUDA_CHECK(cudaFree(d_handles));
    PeerReduceDesc h = {};
    h.world = c->world; h.rank = c->rank; h.cap = cap; h.epoch = 0; h.ticket = 0; h.error = 0;
    for (int q = 0; q < c->world; ++q) {
        uint8_t *base = c->d_mailbox;
        if (q != c->rank) {
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess)
                return fail(B2_ERR_CUDA, "b2_comm_enable_peer_reduce: cannot map the mailbox of rank %d: %s", q, cudaGetErrorString(e));
            c->peer_maps[q] = p;
            base = static_cast<uint8_t *>(p);
        }
        h.flags[q] = reinterpret_cast<uint32_t *>(base);
        h.mail[q] = reinterpret_cast<unsigned long long *>(base + flag_bytes);
    }
    PeerReduceDesc *d_desc = nullptr;
    B2_CUDA_CHECK(cudaMalloc(&d_desc, sizeof(h)));
    cudaError_t ce = cudaMemcpy(d_desc, &h, sizeof(h), cudaMemcpyHostToDevice);
    // nobody writes into a mailbox before every rank has zeroed and mapped: one more (tiny) collective as a barrier
    int64_t *d_one = reinterpret_cast<int64_t *>(c->d_mailbox + flag_bytes);
    ncclResult_t nr = ce == cudaSuccess ? api->AllReduce(d_one, d_one, 1, ncclInt64, ncclSum, c->comm, nullptr) : ncclSuccess;
    if (ce == cudaSuccess && nr == ncclSuccess) ce = cudaStreamSynchronize(nullptr);
    if (ce == cudaSuccess && nr == ncclSuccess) ce = cudaMemset(d_one, 0, 8);
    if (ce != cudaSuccess || nr != ncclSuccess) {
        cudaFree(d_desc);
        return ce != cudaSuccess ? fail(B2_ERR_CUDA, "b2_comm_enable_peer_reduce: %s", cudaGetErrorString(ce))
                                 : fail(B2_ERR_NCCL, "b2_comm_enable_peer_reduce: %s", api->GetErrorString(nr));
    }
    c->d_peer = d_desc;
    c->peer_cap = cap;
    return B2_OK;
}

// 1 = a peer did not arrive within the spin budget of some earlier peer collective (results of that step are invalid).
extern "C" int b2_comm_peer_status(b2_comm *c, int *timed_out) {
    using namespace b2;
    B2_REQUIRE(c != nullptr && timed_out != nullptr && c->d_peer != nullptr, "b2_comm_peer_status: peer reduce is not enabled");
    PeerReduceDesc h;
    B2_CUDA_CHECK(cudaMemcpy(&h, c->d_peer, sizeof(h), cudaMemcpyDeviceToHost));
    *timed_out = int(h.error);
    return B2_OK;
}

extern "C" int b2_peer_allreduce_i64(b2_comm *c, int64_t *d_values, uint32_t count, void *stream) {
    using namespace b2;
    B2_REQUIRE(c != nullptr && c->d_peer != nullptr, "b2_peer_allreduce_i64: call b2_comm_enable_peer_reduce first");
    B2_REQUIRE(count == 0 || d_values != nullptr, "b2_peer_allreduce_i64: null pointer");
    B2_REQUIRE(count <= c->peer_cap, "b2_peer_allreduce_i64: %u values exceed the mailbox capacity %u", count, c->peer_cap);
    if (count == 0) return B2_OK;
    // Same L1 / shared-memory split as the tally kernel it usually runs beside: an SM only changes its carve-out when
    // idle, so a kernel with another preference would keep one SM away from the tally's two-CTAs-per-SM wave.
    static std::once_flag once[64];
    std::call_once(once[c->device & 63], [] {
        cudaFuncSetAttribute(peer_allreduce_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    });
    peer_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(c->d_peer, reinterpret_cast<unsigned long long *>(d_values), count);
    B2_LAUNCH_CHECK("peer_allreduce_kernel");
    return B2_OK;
}

extern "C" int b2_label_tally_reduce(b2_comm *c, const int32_t *d_image_idx, const uint8_t *d_class_idx, const uint8_t *d_active,
                                     uint64_t rows, uint32_t image_base, uint32_t n_images, uint32_t k, uint32_t flags,
                                     int32_t *d_counts, int64_t *d_partials_hist, void *stream) {
    using namespace b2;
    B2_REQUIRE(c != nullptr && d_partials_hist != nullptr, "b2_label_tally_reduce: null pointer");
    const uint32_t n_values = k + B2_PARTIALS_EXTRA + B2_AGREE_BINS;
    if (c->world == 1)
        return b2_label_tally(d_image_idx, d_class_idx, d_active, rows, image_base, n_images, k, flags, d_counts, d_partials_hist,
                              d_partials_hist + k + B2_PARTIALS_EXTRA, stream);
    if (c->d_peer != nullptr && (flags & B2_TALLY_SORTED) && n_values <= c->peer_cap)
        return label_tally_impl(d_image_idx, d_class_idx, d_active, rows, image_base, n_images, k, flags, d_counts, d_partials_hist,
                                d_partials_hist + k + B2_PARTIALS_EXTRA, c->d_peer, stream);
    // any-order rows or no peer mailbox: the tally, then NCCL
    int rc = b2_label_tally(d_image_idx, d_class_idx, d_active, rows, image_base, n_images, k, flags, d_counts, d_partials_hist,
                            d_partials_hist + k + B2_PARTIALS_EXTRA, stream);
    if (rc != B2_OK) return rc;
    return b2_allreduce_i64(c, d_partials_hist, n_values, stream);
}
