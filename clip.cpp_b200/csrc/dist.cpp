// dist.cpp -- NCCL communicators for libclip_b200.so, loaded with dlopen, no torch / MPI dependency (see dist.h).
#include "dist.h"

#include <dlfcn.h>
#include <nccl.h>        // types and prototypes only: every entry point is resolved with dlsym from libnccl.so.2
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <mutex>

namespace cb {

struct NcclApi {
    decltype(&ncclGetVersion) GetVersion;
    decltype(&ncclGetUniqueId) GetUniqueId;
    decltype(&ncclCommInitRank) CommInitRank;
    decltype(&ncclCommInitAll) CommInitAll;
    decltype(&ncclCommDestroy) CommDestroy;
    decltype(&ncclAllGather) AllGather;
    decltype(&ncclAllReduce) AllReduce;
    decltype(&ncclGroupStart) GroupStart;
    decltype(&ncclGroupEnd) GroupEnd;
    decltype(&ncclGetErrorString) GetErrorString;
};

namespace {

std::mutex g_mu;
NcclApi g_api;
bool g_loaded = false;
std::string g_load_err;
const double g_lib_load_time = (double)time(nullptr);
std::atomic<int> g_rdzv_seq{0};

template <class F>
bool sym(void* h, const char* name, F& out, std::string& err) {
    out = reinterpret_cast<F>(dlsym(h, name));
    if (!out) { err = std::string("libnccl: missing symbol ") + name; return false; }
    return true;
}

std::string nerr(const NcclApi* a, const char* what, ncclResult_t r) {
    return std::string(what) + " failed: " + (a && a->GetErrorString ? a->GetErrorString(r) : "?");
}

}  // namespace

const NcclApi* nccl_api(std::string& err) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_loaded) return &g_api;
    if (!g_load_err.empty()) { err = g_load_err; return nullptr; }
    const char* names[] = {getenv("CLIP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (h) break;
    }
    if (!h) { g_load_err = err = std::string("cannot load libnccl.so.2 (multi-GPU needs NCCL): ") + (dlerror() ? dlerror() : ""); return nullptr; }
    std::string e;
    const bool ok = sym(h, "ncclGetVersion", g_api.GetVersion, e) && sym(h, "ncclGetUniqueId", g_api.GetUniqueId, e) &&
                    sym(h, "ncclCommInitRank", g_api.CommInitRank, e) && sym(h, "ncclCommInitAll", g_api.CommInitAll, e) &&
                    sym(h, "ncclCommDestroy", g_api.CommDestroy, e) && sym(h, "ncclAllGather", g_api.AllGather, e) &&
                    sym(h, "ncclAllReduce", g_api.AllReduce, e) && sym(h, "ncclGroupStart", g_api.GroupStart, e) &&
                    sym(h, "ncclGroupEnd", g_api.GroupEnd, e) && sym(h, "ncclGetErrorString", g_api.GetErrorString, e);
    if (!ok) { g_load_err = err = e; return nullptr; }
    g_loaded = true;
    return &g_api;
}

int dist_nccl_version() {
    std::string e;
    const NcclApi* a = nccl_api(e);
    int v = 0;
    if (a) a->GetVersion(&v);
    return v;
}

bool dist_unique_id(void* out128, std::string& err) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    const NcclApi* a = nccl_api(err);
    if (!a) return false;
    ncclUniqueId id;
    const ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) { err = nerr(a, "ncclGetUniqueId", r); return false; }
    memcpy(out128, &id, 128);
    return true;
}

// Rendezvous of the 128-byte id through a file on the node.  Default path: /tmp/clip_b200_rdzv_<MASTER_PORT>_<parent pid>.<seq> --
// every worker of one launcher (torchrun agent, mpirun, a test harness) shares the parent pid and MASTER_PORT, and <seq> counts
// the rendezvous calls of this process, so repeated initialisations do not collide.  Rank 0 writes tmp + rename (atomic); readers
// poll, and ignore files written long before this library was loaded (left over from a crashed run).
static thread_local std::string t_last_rdzv_path;

// after ncclCommInitRank has returned on rank 0 every rank has read the id (the call is collective): the file can go
void dist_rendezvous_done(int rank) {
    if (rank == 0 && !t_last_rdzv_path.empty()) unlink(t_last_rdzv_path.c_str());
    t_last_rdzv_path.clear();
}

bool dist_rendezvous_id(int rank, int world, const char* rendezvous, void* id128, std::string& err) {
    (void)world;
    char path[512];
    const int seq = g_rdzv_seq.fetch_add(1);
    if (rendezvous && *rendezvous) snprintf(path, sizeof path, "%s.%d", rendezvous, seq);
    else {
        const char* port = getenv("MASTER_PORT");
        const char* dir = getenv("CLIP_B200_RDZV_DIR");
        snprintf(path, sizeof path, "%s/clip_b200_rdzv_%s_%d.%d", dir ? dir : "/tmp", port ? port : "0", (int)getppid(), seq);
    }
    struct Rec { unsigned char id[128]; double t; } rec;
    if (rank == 0) {
        if (!dist_unique_id(rec.id, err)) return false;
        rec.t = (double)time(nullptr);
        char tmp[600];
        snprintf(tmp, sizeof tmp, "%s.tmp.%d", path, (int)getpid());
        unlink(path);
        FILE* f = fopen(tmp, "wb");
        if (!f || fwrite(&rec, sizeof rec, 1, f) != 1) { if (f) fclose(f); err = std::string("cannot write rendezvous file ") + tmp; return false; }
        fclose(f);
        if (rename(tmp, path) != 0) { err = std::string("cannot publish rendezvous file ") + path; return false; }
        memcpy(id128, rec.id, 128);
        t_last_rdzv_path = path;
        return true;
    }
    const double deadline = (double)time(nullptr) + 600.0;
    for (;;) {
        FILE* f = fopen(path, "rb");
        if (f) {
            const size_t n = fread(&rec, sizeof rec, 1, f);
            fclose(f);
            if (n == 1 && rec.t >= g_lib_load_time - 120.0) { memcpy(id128, rec.id, 128); return true; }
        }
        if ((double)time(nullptr) > deadline) { err = std::string("timed out waiting for rank 0's rendezvous file ") + path; return false; }
        usleep(20 * 1000);
    }
}

bool dist_init_rank(DistComm& dc, int rank, int world, const void* id128, std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a) return false;
    if (world < 1 || rank < 0 || rank >= world) { err = "bad rank / world size"; return false; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    const ncclResult_t r = a->CommInitRank(&comm, world, id, rank);      // binds to the CURRENT CUDA device
    if (r != ncclSuccess) { err = nerr(a, "ncclCommInitRank", r); return false; }
    dc.comm = comm; dc.rank = rank; dc.world = world;
    return true;
}

bool dist_init_all(DistComm* comms, const int* devices, int n, std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a) return false;
    ncclComm_t tmp[64];
    if (n < 1 || n > 64) { err = "bad device count"; return false; }
    const ncclResult_t r = a->CommInitAll(tmp, n, devices);
    if (r != ncclSuccess) { err = nerr(a, "ncclCommInitAll", r); return false; }
    for (int i = 0; i < n; i++) { comms[i].comm = tmp[i]; comms[i].rank = i; comms[i].world = n; }
    return true;
}

void dist_destroy(DistComm& dc) {
    if (!dc.comm) return;
    std::string e;
    const NcclApi* a = nccl_api(e);
    if (a) a->CommDestroy((ncclComm_t)dc.comm);
    dc.comm = nullptr; dc.rank = 0; dc.world = 1;
}

bool dist_all_gather(const DistComm& dc, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st, std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a || !dc.comm) { if (a) err = "no communicator (call clip_b200_dist_init first)"; return false; }
    const ncclResult_t r = a->AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)dc.comm, st);
    if (r != ncclSuccess) { err = nerr(a, "ncclAllGather", r); return false; }
    return true;
}

bool dist_all_reduce_max_f64(const DistComm& dc, double* d_buf, size_t n, cudaStream_t st, std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a || !dc.comm) { if (a) err = "no communicator"; return false; }
    const ncclResult_t r = a->AllReduce(d_buf, d_buf, n, ncclDouble, ncclMax, (ncclComm_t)dc.comm, st);
    if (r != ncclSuccess) { err = nerr(a, "ncclAllReduce", r); return false; }
    return true;
}

bool dist_all_reduce_sum_i32(const DistComm& dc, int32_t* d_buf, size_t n, cudaStream_t st, std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a || !dc.comm) { if (a) err = "no communicator"; return false; }
    const ncclResult_t r = a->AllReduce(d_buf, d_buf, n, ncclInt32, ncclSum, (ncclComm_t)dc.comm, st);
    if (r != ncclSuccess) { err = nerr(a, "ncclAllReduce", r); return false; }
    return true;
}

bool dist_group_start(std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a) return false;
    const ncclResult_t r = a->GroupStart();
    if (r != ncclSuccess) { err = nerr(a, "ncclGroupStart", r); return false; }
    return true;
}
bool dist_group_end(std::string& err) {
    const NcclApi* a = nccl_api(err);
    if (!a) return false;
    const ncclResult_t r = a->GroupEnd();
    if (r != ncclSuccess) { err = nerr(a, "ncclGroupEnd", r); return false; }
    return true;
}

}  // namespace cb
