// Host runtime: per-device streams, workspace pool, NCCL communicator, and the executors that turn
// a plan into kernel launches.
//
// Reference counterparts: one persistent stream per device (platform.h:456-467), buffer pool
// (alloc.h:323-391), causalConnectTo (platform.h:371-409), send_receive (dist.h:1426-1573).
#pragma once
#include "kernels.hpp"
#include "plan.hpp"
#include <set>
#include <vector>

namespace sbb {

    struct DeviceState {
        int id = -1;
        cudaStream_t stream = nullptr;      ///< all kernels of the library for this device
        cudaStream_t comm_stream = nullptr; ///< NCCL traffic, overlapped with `stream`
        cudaStream_t aux_stream = nullptr;  ///< unpack / local kernels that overlap the pack kernels of an exchange
        cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr;
    };

    DeviceState &device_state(int device);
    void use_device(int device);
    void enable_peer(int a, int b);
    void destroy_all_streams();
    /// Make every listed device's stream wait for the work already queued on the others
    void order_streams(const std::set<int> &devs);

    /// Cached device allocations (reused by later calls on the same stream)
    void *pool_alloc(int device, size_t bytes);
    void pool_free(int device, void *p);
    void pool_clear();

    struct Comm {
        void *nccl = nullptr; ///< ncclComm_t
        int nranks = 1, rank = 0, device = 0;
        // Peer-memory transport: every rank owns a receive arena (two halves, used alternately) that
        // all other ranks map with CUDA IPC; pack kernels store straight into the receiver's arena
        // over NVLink and one small NCCL all-reduce per exchange is the barrier.
        bool p2p = false;          ///< transport usable (decided collectively at creation)
        char *arena = nullptr;     ///< my arena
        size_t half_bytes = 0;     ///< size of one half
        std::vector<char *> peer;  ///< every rank's arena as mapped here (peer[rank] == arena)
        unsigned long long epoch = 0;
        int *flag = nullptr;       ///< device scratch for the barrier / handle exchange
        // Flag-based signalling (replaces the NCCL all-reduce barrier on the data path): every
        // arena ends with one 64-bit slot per rank; a sender raises its slot in every receiver's
        // arena to the sequence number of the round once its pack kernels are done.
        bool signal = true;                      ///< SBB_P2P_SIGNAL=0 selects the NCCL barrier instead
        unsigned long long seq = 0;              ///< rounds signalled so far (all ranks agree)
        unsigned long long *flags = nullptr;     ///< my slots (inside my arena allocation)
        unsigned long long **peer_flags = nullptr; ///< device array: every rank's slots as mapped here
        // Failure handling: a wait kernel that gives up (SBB_WAIT_TIMEOUT_S, default 60 s) writes the
        // missing rank + 1 here (pinned host memory mapped into the device); an exception inside an
        // exchange poisons the communicator, because the ranks no longer agree on epoch / seq.
        int *error_host = nullptr, *error_dev = nullptr;
        bool poisoned = false;
    };

    /// Throws when the communicator is poisoned or one of its wait kernels timed out
    void comm_check(Comm *c);

    void nccl_unique_id(void *id128);
    Comm *comm_create(const void *id128, int nranks, int rank, int device);
    void comm_destroy(Comm *c);

    /// A component buffer as given by the caller
    struct Buffer {
        void *ptr = nullptr;
        bool host = false;
        int device = 0;
    };

    void execute_copy(const CopyPlan &plan, const CopyArgs &args, int dtype0, int dtype1,
                      const double *alpha, const std::vector<Buffer> &v0,
                      const std::vector<Buffer> &v1, Comm *comm,
                      const std::vector<Buffer> *mask_a = nullptr,
                      const std::vector<Buffer> *mask_b = nullptr);

    long long launch_count(bool reset);

    /// Granularity of the pipelined exchange (bytes per peer and round); SBB_CHUNK_MB overrides
    int64_t exchange_chunk_bytes();

    /// Default device for staging host buffers when no GPU component takes part
    int default_device(Comm *comm);

} // namespace sbb
