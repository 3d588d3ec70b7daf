// Copy planner: turns "copy the range (from0,size0) of the partitioned tensor v0 into v1 at from1"
// into a list of strided-box operations for one rank.
//
// What the reference does here: get_indices_to_send / get_indices_to_receive (dist.h:1787,:1851),
// has_full_support (dist.h:666), may_need_communications (dist.h:2158), prepare_pack / pack_component
// (dist.h:749,:878), prepare_unpack (dist.h:1152) and the per-box `local_copy` loop of copy_request
// (dist.h:2392-2435).  This planner is written from the semantics (oracle/oracle.py),
// not from that code: all geometry happens in range space (geometry.hpp) and the result is a flat
// list of box operations with explicit strides, so no index vector is ever built.
#pragma once
#include "geometry.hpp"
#include <memory>
#include <string>

namespace sbb {

    /// One strided-box operation. Dimensions are listed in destination order, fastest first.
    struct BoxOp {
        enum Kind {
            Local,  ///< src part -> dst part, both on this rank
            Pack,   ///< src part (this rank) -> send buffer of `peer`
            Unpack, ///< receive buffer of `peer` -> dst part (this rank)
            Zero    ///< zero-fill a box of a local dst part
        } kind;
        int src_comp = -1, dst_comp = -1; ///< local component indices (or -1)
        int src_part = -1, dst_part = -1; ///< global part indices rank*ncomponents+comp
        int peer = -1;                    ///< other rank for Pack/Unpack
        Coor size;                        ///< extents
        std::vector<int64_t> sstride, dstride;
        int64_t soff = 0, doff = 0; ///< element offsets (for Pack/Unpack: inside the peer's segment)
        int rot = 0; ///< rotation of the first listed dimension in the destination (see sbk_box_desc)
        int64_t volume() const { return sbb::volume(size); }
    };

    struct CopyPlan {
        std::vector<BoxOp> ops;
        int nranks = 1, rank = 0;
        /// Elements of Q exchanged with each rank (index = peer rank)
        std::vector<int64_t> send_elems, recv_elems;
        bool needs_comm = false; ///< this rank sends or receives something
        bool any_comm = false;   ///< some rank does (all ranks agree on this one)
        /// Peer-memory transport: element offset of my segment inside receiver q's arena, of sender
        /// r's segment inside mine, and the largest arena any rank needs (all ranks agree)
        std::vector<int64_t> send_seg_off, recv_seg_off;
        int64_t arena_elems = 0;
        int64_t max_pair_elems = 0; ///< longest message of the whole exchange (all ranks agree)
        /// Phase of my message to every peer: a proper colouring of the edges of the exchange
        /// (sender -> receiver pairs), computed identically on every rank, so that senders that
        /// visit their receivers in phase order do not meet at a receiver (index = peer rank)
        std::vector<int> send_phase;
        int nphases = 0; ///< phases of the whole exchange (all ranks agree)
        std::string describe() const;
    };

    struct CopyArgs {
        int nd0 = 0, nd1 = 0;
        std::vector<Box> p0, p1; ///< all parts of all ranks, [rank][component] flattened
        int ncomp0 = 1, ncomp1 = 1;
        std::string o0, o1;
        Coor from0, size0, dim0, from1, dim1;
        int nranks = 1, rank = 0;
        int co = FastToSlow;
        bool add = false;
        bool alpha_is_zero = false;
        /// Alignment, in elements of Q, of every box inside a message segment
        int wire_align = 1;
        /// Remote boxes larger than this many bytes are cut in pieces (0 = never); see plan.cpp
        int64_t chunk_bytes = 0;
    };

    /// Validate labels and sizes the way the reference does (toArray tensor.h:266, check_isomorphic
    /// tensor.h:495, "Invalid copy operation" dist.h:2293); throws std::runtime_error.
    void check_copy_args(const CopyArgs &a);

    /// Build the plan for a.rank. Deterministic: every rank derives consistent message layouts.
    std::shared_ptr<const CopyPlan> make_copy_plan(const CopyArgs &a);

    /// Cached version (key = every field of CopyArgs)
    std::shared_ptr<const CopyPlan> get_copy_plan(const CopyArgs &a);
    void clear_plan_cache();
    /// Plans held by the cache
    size_t plan_cache_size();

} // namespace sbb
