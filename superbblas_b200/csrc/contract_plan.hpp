// Contraction driver: label classification, owner-computes decomposition over the partition of the
// larger operand, optional re-partition of the other operand, local fused kernels, reduction of the
// partial results into the caller's output partition.
//
// Reference counterparts: contraction_normalized (dist.h:3092-3196), get_partitions_for_contraction
// (dist.h:3039-3090), remove_repetitions (dist.h:3001-3028), suggested_orders_for_contraction
// (tensor.h:1271-1429; only its label classification and error rules survive here, the re-ordering
// is unnecessary because the kernels take arbitrary strides).
#pragma once
#include "runtime.hpp"

namespace sbb {

    struct TensorArg {
        int nd = 0;
        std::string o;
        int ncomp = 1;
        std::vector<Box> p; ///< all parts of all ranks
        Coor from, size, dim;
    };

    struct ContractionArgs {
        int dtype = SBB_C128;
        double alpha[2] = {1, 0}, beta[2] = {0, 0};
        TensorArg t0, t1, tr;
        bool conj0 = false, conj1 = false;
        int co = FastToSlow, nranks = 1, rank = 0;
    };

    void execute_contraction(const ContractionArgs &a, const std::vector<Buffer> &v0,
                             const std::vector<Buffer> &v1, const std::vector<Buffer> &vr,
                             Comm *comm);

} // namespace sbb
