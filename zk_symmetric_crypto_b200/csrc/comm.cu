// NCCL through dlopen (see comm.hpp).
#include <dlfcn.h>
#include <string.h>
#include <mutex>
#include <stdexcept>
#include <string>
#include "comm.hpp"
#if __has_include(<nccl.h>)
#include <nccl.h>  // enum values only; every NCCL function is resolved with dlsym below
#define S2C_HAVE_NCCL_H 1
#endif

namespace {
struct UniqueId { char internal[128]; };
typedef int (*fn_get_id)(UniqueId*);
typedef int (*fn_init_rank)(void**, int, UniqueId, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_void)();
typedef int (*fn_sendrecv)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);

struct Api {
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_void group_start = nullptr, group_end = nullptr;
    fn_sendrecv send = nullptr, recv = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_errstr errstr = nullptr;
};

Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) throw std::runtime_error(std::string("NCCL not available: ") + dlerror());
        a.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
        a.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
        a.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
        a.group_start = (fn_void)dlsym(h, "ncclGroupStart");
        a.group_end = (fn_void)dlsym(h, "ncclGroupEnd");
        a.send = (fn_sendrecv)dlsym(h, "ncclSend");
        a.recv = (fn_sendrecv)dlsym(h, "ncclRecv");
        a.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
        a.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
        if (!a.get_id || !a.init_rank || !a.destroy || !a.group_start || !a.group_end || !a.send || !a.recv || !a.allreduce)
            throw std::runtime_error("NCCL symbols missing");
    });
    return a;
}

void ck(int rc, const char* what) {
    if (rc != 0) {
        const char* s = api().errstr ? api().errstr(rc) : "?";
        throw std::runtime_error(std::string(what) + ": NCCL error " + std::to_string(rc) + " (" + s + ")");
    }
}
#ifdef S2C_HAVE_NCCL_H
constexpr int kUint32 = (int)ncclUint32, kInt32 = (int)ncclInt32, kMin = (int)ncclMin, kSum = (int)ncclSum;
#else
constexpr int kUint32 = 3, kInt32 = 2, kMin = 3, kSum = 0;  // ncclUint32, ncclInt32, ncclMin (nccl.h of NCCL 2.x)
#endif
}  // namespace

void comm_unique_id(uint8_t out[128]) {
    UniqueId id;
    ck(api().get_id(&id), "ncclGetUniqueId");
    memcpy(out, id.internal, 128);
}
void comm_init(Comm& c, int rank, int world, const uint8_t idb[128]) {
    UniqueId id;
    memcpy(id.internal, idb, 128);
    ck(api().init_rank(&c.comm, world, id, rank), "ncclCommInitRank");
    c.rank = rank;
    c.world = world;
}
void comm_destroy(Comm& c) {
    if (c.comm) api().destroy(c.comm);
    if (c.scratch) cudaFree(c.scratch);
    c.scratch = nullptr;
    c.comm = nullptr;
    c.rank = 0;
    c.world = 1;
}
void comm_group_start() { ck(api().group_start(), "ncclGroupStart"); }
void comm_group_end() { ck(api().group_end(), "ncclGroupEnd"); }
void comm_send_u32(const Comm& c, const uint32_t* p, size_t words, int peer, cudaStream_t st) {
    ck(api().send((void*)p, words, kUint32, peer, c.comm, st), "ncclSend");
}
void comm_recv_u32(const Comm& c, uint32_t* p, size_t words, int peer, cudaStream_t st) {
    ck(api().recv((void*)p, words, kUint32, peer, c.comm, st), "ncclRecv");
}
// all ranks' `words`-word blocks, rank-major, into recv_dev (grouped send/recv; the local block is copied)
void comm_allgather_u32(const Comm& c, const uint32_t* send_dev, uint32_t* recv_dev, size_t words, cudaStream_t st) {
    comm_group_start();
    for (int r = 0; r < c.world; r++) {
        if (r == c.rank) {
            cudaMemcpyAsync(recv_dev + (size_t)r * words, send_dev, words * 4, cudaMemcpyDeviceToDevice, st);
        } else {
            comm_send_u32(c, send_dev, words, r, st);
            comm_recv_u32(c, recv_dev + (size_t)r * words, words, r, st);
        }
    }
    comm_group_end();
}
void comm_bcast_u32(const Comm& c, uint32_t* buf_dev, size_t words, int root, cudaStream_t st) {
    if (!c.active()) return;
    comm_group_start();
    if (c.rank == root) {
        for (int r = 0; r < c.world; r++)
            if (r != root) comm_send_u32(c, buf_dev, words, r, st);
    } else {
        comm_recv_u32(c, buf_dev, words, root, st);
    }
    comm_group_end();
}
// stream-ordered barrier: a one-word all-reduce on the communicator's scratch word.  Work enqueued on `st` after it starts
// only when every rank's work enqueued before it has finished (including stores it made into peer memory).
void comm_barrier(Comm& c, cudaStream_t st) {
    if (!c.active()) return;
    if (!c.scratch) {
        if (cudaMalloc(&c.scratch, 64) != cudaSuccess) throw std::runtime_error("cudaMalloc (comm scratch)");
        cudaMemsetAsync(c.scratch, 0, 64, st);
    }
    ck(api().allreduce(c.scratch, c.scratch, 1, kInt32, kMin, c.comm, st), "ncclAllReduce (barrier)");
}
// in-place sum of u32 words over the ranks (used to merge partial results with disjoint support, so no carry or modular
// reduction is involved)
void comm_allreduce_sum_u32(const Comm& c, uint32_t* buf_dev, size_t words, cudaStream_t st) {
    if (!c.active()) return;
    ck(api().allreduce(buf_dev, buf_dev, words, kUint32, kSum, c.comm, st), "ncclAllReduce (sum)");
}
int comm_min_int(const Comm& c, int v, cudaStream_t st) {
    if (!c.active()) return v;
    struct DevInt {  // freed on every path, including a throwing ck()
        int* p = nullptr;
        DevInt() { if (cudaMalloc(&p, sizeof(int)) != cudaSuccess) throw std::runtime_error("cudaMalloc"); }
        ~DevInt() { cudaFree(p); }
    } d;
    cudaMemcpyAsync(d.p, &v, sizeof(int), cudaMemcpyHostToDevice, st);
    ck(api().allreduce(d.p, d.p, 1, kInt32, kMin, c.comm, st), "ncclAllReduce");
    cudaMemcpyAsync(&v, d.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    return v;
}
