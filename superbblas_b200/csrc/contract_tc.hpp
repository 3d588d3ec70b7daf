// Blackwell tensor-core path of the contraction for complex float operands
// (kernels_contract_tc.cu): TMA tile loads -> 3xTF32 split in shared memory -> tcgen05.mma with the
// accumulators in tensor memory.  Interface between the dispatcher (kernels_contract.cu) and the
// kernel; see the .cu file for the design.
#pragma once
#include "kernels.hpp"

namespace sbb {
    namespace tc {

        struct Dim {
            int size;
            long long s0, s1, sr; ///< strides (elements) in v0, v1, vr
        };

        /// vr[t,m,n] = alpha * sum_k f0(v0[t,m,k]) f1(v1[t,n,k]) + beta * vr[t,m,n] with the
        /// contracted index contiguous (stride 1) in both operands
        struct Problem {
            int nT;
            Dim T[2]; ///< batch dims, first fastest
            Dim M, N; ///< the free index of v0 / v1 (size 1 when absent)
            long long K;
            int conj0, conj1;
        };

        /// Whether the operands can be described to the TMA unit (16-byte aligned bases and strides)
        bool eligible(const Problem &p, const void *v0, const void *v1);

        /// Complex float.  If `describe` is given nothing is launched.
        void launch_c64(const Problem &p, const double *alpha, const void *v0, const void *v1,
                        const double *beta, void *vr, int device, cudaStream_t stream,
                        std::string *describe);

    } // namespace tc
} // namespace sbb
