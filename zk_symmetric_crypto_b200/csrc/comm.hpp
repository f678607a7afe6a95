// Minimal NCCL binding for the row-sharded single-proof mode (BASELINE cfg-5): the LDE tiles of one large trace
// are transformed column-sharded (each rank transforms whole columns) and exchanged with a grouped send/recv all-to-all so
// that every rank holds a row range of ALL columns for leaf hashing and constraint evaluation.
// libnccl.so.2 is resolved at run time (dlopen; the copy PyTorch already loaded is reused), so the library builds and loads
// without NCCL and the single-GPU paths never touch it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

struct Comm {
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, world = 1;
    int* scratch = nullptr;  // device word for comm_barrier
    bool active() const { return world > 1; }
};

void comm_unique_id(uint8_t out[128]);
void comm_init(Comm& c, int rank, int world, const uint8_t id[128]);
void comm_destroy(Comm& c);
void comm_group_start();
void comm_group_end();
void comm_send_u32(const Comm& c, const uint32_t* p, size_t words, int peer, cudaStream_t st);
void comm_recv_u32(const Comm& c, uint32_t* p, size_t words, int peer, cudaStream_t st);
void comm_allgather_u32(const Comm& c, const uint32_t* send_dev, uint32_t* recv_dev, size_t words, cudaStream_t st);
void comm_bcast_u32(const Comm& c, uint32_t* buf_dev, size_t words, int root, cudaStream_t st);
void comm_barrier(Comm& c, cudaStream_t st);
void comm_allreduce_sum_u32(const Comm& c, uint32_t* buf_dev, size_t words, cudaStream_t st);
// minimum over ranks of a host int (synchronises `st`)
int comm_min_int(const Comm& c, int v, cudaStream_t st);
