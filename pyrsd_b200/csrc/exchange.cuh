// exchange.cuh -- the per-step exchange of a walker ensemble sharded over the GPUs of one box.  Every rank's epilogue
// kernel writes its per-walker results straight into its own slot of its receive buffer; a push kernel stores that slot
// into the same slot of every peer's receive buffer over NVLink / NVSwitch peer memory (CUDA IPC mappings between the
// one-process-per-GPU ranks) and a flag exchange follows -- no staging buffer, no library collective (SURVEY 8e:
// "assembly kernel writes directly into the ... buffer").
#pragma once
#include "common.h"

#include <vector>

namespace prsd {

static const int EXCH_MAX_RANKS = 16;
static const int EXCH_HANDLE_BYTES = 128;          // two cudaIpcMemHandle_t: receive buffers, flags

struct PeerTable {                                  // passed to kernels by value
    int world, rank;
    double* recv[EXCH_MAX_RANKS];                   // this rank's slot in peer p's receive buffer of the current parity
};

struct Exchange {
    Context* ctx;
    int world, rank;
    size_t n_per_rank;                              // doubles every rank contributes per step
    double* recv_local = nullptr;                   // [2][world][n_per_rank]  (cudaMalloc: whole allocations are shared)
    unsigned long long* flags_local = nullptr;      // [world]: flags_local[p] = last step rank p has delivered here
    int* d_status = nullptr;                        // 0 ok, 1 a wait timed out
    double* recv_peer[EXCH_MAX_RANKS] = {nullptr};
    unsigned long long* flags_peer[EXCH_MAX_RANKS] = {nullptr};
    bool opened = false;
    unsigned long long step = 0;

    Exchange(Context& c, int world_, int rank_, size_t n_per_rank_);
    ~Exchange();
    void handles(unsigned char* out128) const;
    void open(const unsigned char* all /*[world][128]*/);
    // start the NEXT step: returns where the producing kernel writes this rank's n_per_rank results (its own slot)
    double* begin_step();
    // after the producing kernel on ctx->stream: push the slot to every peer, raise this rank's flag there, wait for theirs
    void finish_step();
    const double* received() const { return recv_local + (size_t)(step & 1)*world*n_per_rank; }
    int status();
};

} // namespace prsd
