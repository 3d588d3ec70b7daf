// TEST INFRASTRUCTURE — not product code.
//
// C-ABI shim around the UNMODIFIED reference headers under /root/reference/include (never copied
// into this repo).  It is compiled in place by oracle/Makefile into oracle/_ref/libsbref.so and is
// used only (a) by tests/ to pin the numpy oracle and to cross-check the CUDA path, (b) by
// bench.py's `cpu_baseline` / `--impl reference` legs.  The product never loads it.
//
// The entry points mirror include/superbblas_b200.h one-to-one (same argument order and meaning)
// so that a test can call the reference and the product with identical inputs.
//
// Reference interfaces wrapped:
//   superbblas::copy               include/superbblas/dist.h:3583  (no-MPI overload)
//   superbblas::contraction        include/superbblas/dist.h:3701  (no-MPI overload) via
//       detail::contraction_normalized (dist.h:3092), which is the runtime-rank form the public
//       template lowers to (dist.h:3220-3245)
//   superbblas::basic_partitioning dist.h:3393 and :3477
//   superbblas::partitioning_distributed_procs dist.h:3318
//   superbblas::make_hole          dist.h:3802
//
// Runtime rank: the reference is templated on the number of dimensions.  Exactly like the
// reference's own `dummy_normalize_copy` (dist.h:2472) we pad every tensor to SBREF_ND dimensions
// with size-1 dimensions carrying fresh labels; this does not change any result.

#include "superbblas.h"
#include <complex>
#include <cstring>
#include <string>
#include <vector>

using namespace superbblas;

#ifndef SBREF_ND
#    define SBREF_ND 8
#endif

static thread_local std::string g_err;

namespace {
    constexpr std::size_t ND = SBREF_ND;

    struct Padded {
        std::vector<PartitionItem<ND>> p;
        Coor<ND> from, size, dim;
        char o[ND + 1];
    };

    // Pad a rank-nd description to ND dims: extra dims have from=0,size=dim=1 and unused labels
    Padded pad(int nd, const int *p, int nparts, const char *o, const int *from, const int *size,
               const int *dim, std::vector<bool> &used) {
        if (nd > (int)ND) throw std::runtime_error("sbref: too many dimensions");
        if ((int)std::strlen(o) != nd) throw std::runtime_error("sbref: order length mismatch");
        Padded r;
        r.p.resize(nparts);
        for (int i = 0; i < nparts; ++i) {
            for (int k = 0; k < (int)ND; ++k) {
                r.p[i][0][k] = k < nd ? p[(i * 2 + 0) * nd + k] : 0;
                r.p[i][1][k] = k < nd ? p[(i * 2 + 1) * nd + k] : 1;
            }
            // The reference normalises empty ranges to all zeros
            bool empty = false;
            for (int k = 0; k < nd; ++k)
                if (r.p[i][1][k] == 0) empty = true;
            if (empty)
                for (int k = 0; k < (int)ND; ++k) r.p[i][0][k] = r.p[i][1][k] = 0;
        }
        for (int k = 0; k < (int)ND; ++k) {
            r.from[k] = k < nd ? from[k] : 0;
            r.size[k] = k < nd ? (size ? size[k] : 0) : 1;
            r.dim[k] = k < nd ? dim[k] : 1;
        }
        std::memcpy(r.o, o, nd);
        int j = nd;
        for (int c = 1; c < 127 && j < (int)ND; ++c)
            if (!used[c]) r.o[j++] = (char)c, used[c] = true;
        r.o[ND] = 0;
        return r;
    }

    void mark(const char *o, std::vector<bool> &used) {
        for (; *o; ++o) used[(int)(unsigned char)*o] = true;
    }

    template <typename T> T mk(const double *a);
    template <> float mk<float>(const double *a) { return (float)a[0]; }
    template <> double mk<double>(const double *a) { return a[0]; }
    template <> int mk<int>(const double *a) { return (int)a[0]; }
    template <> std::complex<float> mk<std::complex<float>>(const double *a) {
        return {(float)a[0], (float)a[1]};
    }
    template <> std::complex<double> mk<std::complex<double>>(const double *a) {
        return {a[0], a[1]};
    }

    template <typename T, typename Q>
    void do_copy(const double *alpha, int nd0, const int *p0, int ncomp0, const char *o0,
                 const int *from0, const int *size0, const int *dim0, const void **v0, int nd1,
                 const int *p1, int ncomp1, const char *o1, const int *from1, const int *dim1,
                 void **v1, int co, int copyadd, const float **mask0, const float **mask1) {
        std::vector<bool> used(128, false);
        mark(o0, used);
        mark(o1, used);
        Padded t0 = pad(nd0, p0, ncomp0, o0, from0, size0, dim0, used);
        Padded t1 = pad(nd1, p1, ncomp1, o1, from1, nullptr, dim1, used);
        std::vector<Context> ctx0(ncomp0, createCpuContext()), ctx1(ncomp1, createCpuContext());
        copy<ND, ND, T, Q>(mk<T>(alpha), t0.p.data(), ncomp0, t0.o, t0.from, t0.size, t0.dim,
                           (const T **)v0, mask0, ctx0.data(), t1.p.data(), ncomp1, t1.o,
                           t1.from, t1.dim, (Q **)v1, mask1, ctx1.data(),
                           co == 0 ? SlowToFast : FastToSlow, copyadd == 0 ? Copy : Add);
    }

    template <typename T>
    void do_contraction(const double *alpha, int nd0, const int *p0, const int *from0,
                        const int *size0, const int *dim0, int ncomp0, const char *o0, int conj0,
                        const void **v0, int nd1, const int *p1, const int *from1,
                        const int *size1, const int *dim1, int ncomp1, const char *o1, int conj1,
                        const void **v1, const double *beta, int ndo, const int *pr,
                        const int *fromr, const int *sizer, const int *dimr, int ncompr,
                        const char *o_r, void **vr, int co) {
        std::vector<bool> used(128, false);
        mark(o0, used);
        mark(o1, used);
        mark(o_r, used);
        Padded t0 = pad(nd0, p0, ncomp0, o0, from0, size0, dim0, used);
        Padded t1 = pad(nd1, p1, ncomp1, o1, from1, size1, dim1, used);
        Padded tr = pad(ndo, pr, ncompr, o_r, fromr, sizer, dimr, used);
        std::vector<Context> ctx0(ncomp0, createCpuContext()), ctx1(ncomp1, createCpuContext()),
            ctxr(ncompr, createCpuContext());
        detail::SelfComm comm = detail::get_comm();
        wait(detail::contraction_normalized<ND, T>(
            mk<T>(alpha), detail::get_from_size(t0.p.data(), ncomp0, comm), t0.from, t0.size,
            t0.dim, detail::toArray<ND>(t0.o, "o0"), conj0 != 0,
            detail::get_components<ND>((T **)v0, nullptr, ctx0.data(), ncomp0, t0.p.data(), comm,
                                       0),
            (std::size_t)nd0, detail::get_from_size(t1.p.data(), ncomp1, comm), t1.from, t1.size,
            t1.dim, detail::toArray<ND>(t1.o, "o1"), conj1 != 0,
            detail::get_components<ND>((T **)v1, nullptr, ctx1.data(), ncomp1, t1.p.data(), comm,
                                       0),
            (std::size_t)nd1, mk<T>(beta), detail::get_from_size(tr.p.data(), ncompr, comm),
            tr.from, tr.size, tr.dim, detail::toArray<ND>(tr.o, "o_r"),
            detail::get_components<ND>((T **)vr, nullptr, ctxr.data(), ncompr, tr.p.data(), comm,
                                       0),
            (std::size_t)ndo, comm, co == 0 ? SlowToFast : FastToSlow));
    }
}

#ifdef SBREF_PART_MAIN
namespace {
    template <std::size_t N> Coor<N> toCoor(const int *v) {
        Coor<N> r;
        for (std::size_t i = 0; i < N; ++i) r[i] = v[i];
        return r;
    }
    template <std::size_t N> void store(const std::vector<PartitionItem<N>> &p, int *out) {
        for (std::size_t i = 0; i < p.size(); ++i)
            for (int j = 0; j < 2; ++j)
                for (std::size_t k = 0; k < N; ++k) out[(i * 2 + j) * N + k] = p[i][j][k];
    }
    template <std::size_t N>
    void bp(const char *order, const int *dim, const int *procs, const char *dist_labels,
            int nprocs, int ncomponents, int *out) {
        store<N>(basic_partitioning<N>(order, toCoor<N>(dim), toCoor<N>(procs), dist_labels, nprocs,
                                       ncomponents),
                 out);
    }
    template <std::size_t N>
    void bpe(const int *dim, const int *procs, int nprocs, int replicate, const int *ext_power,
             int *out) {
        store<N>(basic_partitioning<N>(toCoor<N>(dim), toCoor<N>(procs), nprocs, replicate != 0,
                                       toCoor<N>(ext_power)),
                 out);
    }
    template <std::size_t N>
    void pdp(const char *order, const int *dim, const char *dist_labels, int nprocs, int *out) {
        auto r = partitioning_distributed_procs<N>(order, toCoor<N>(dim), dist_labels, nprocs);
        for (std::size_t k = 0; k < N; ++k) out[k] = r[k];
    }
    template <std::size_t N>
    void mh(const int *from, const int *size, const int *hfrom, const int *hsize, const int *dim,
            int *out, int *nout) {
        auto r = make_hole<N>(toCoor<N>(from), toCoor<N>(size), toCoor<N>(hfrom), toCoor<N>(hsize),
                              toCoor<N>(dim));
        *nout = (int)r.size();
        store<N>(r, out);
    }
}

#endif

#define SBREF_TRY(...)                                                                             \
    try {                                                                                          \
        __VA_ARGS__;                                                                               \
        return 0;                                                                                  \
    } catch (const std::exception &e) {                                                            \
        g_err = e.what();                                                                          \
        return 1;                                                                                  \
    }

// dtype codes (same as include/superbblas_b200.h): 0=f32 1=f64 2=c64 3=c128 4=i32
#define CF std::complex<float>
#define CD std::complex<double>

extern "C" {

#ifdef SBREF_PART_COPY
int SBREF_COPY_NAME(const double *alpha, int nd0, const int *p0, int ncomp0, const char *o0,
                    const int *from0, const int *size0, const int *dim0, const void **v0, int nd1,
                    const int *p1, int ncomp1, const char *o1, const int *from1, const int *dim1,
                    void **v1, int co, int copyadd, const float **mask0, const float **mask1) {
    SBREF_TRY((do_copy<SBREF_T, SBREF_Q>(alpha, nd0, p0, ncomp0, o0, from0, size0, dim0, v0, nd1,
                                         p1, ncomp1, o1, from1, dim1, v1, co, copyadd, mask0,
                                         mask1)));
}
#endif

#ifdef SBREF_PART_CONTRACTION
int SBREF_CONTRACTION_NAME(const double *alpha, int nd0, const int *p0, const int *from0,
                           const int *size0, const int *dim0, int ncomp0, const char *o0, int conj0,
                           const void **v0, int nd1, const int *p1, const int *from1,
                           const int *size1, const int *dim1, int ncomp1, const char *o1, int conj1,
                           const void **v1, const double *beta, int ndo, const int *pr,
                           const int *fromr, const int *sizer, const int *dimr, int ncompr,
                           const char *o_r, void **vr, int co) {
    SBREF_TRY((do_contraction<SBREF_T>(alpha, nd0, p0, from0, size0, dim0, ncomp0, o0, conj0, v0,
                                       nd1, p1, from1, size1, dim1, ncomp1, o1, conj1, v1, beta,
                                       ndo, pr, fromr, sizer, dimr, ncompr, o_r, vr, co)));
}
#endif

#ifdef SBREF_PART_MAIN
const char *sbref_last_error() { return g_err.c_str(); }

// Per-type entry points defined in the other translation units
#    define DECL_COPY(N)                                                                           \
        int N(const double *, int, const int *, int, const char *, const int *, const int *,      \
              const int *, const void **, int, const int *, int, const char *, const int *,       \
              const int *, void **, int, int, const float **, const float **);
DECL_COPY(sbref_copy_0_0)
DECL_COPY(sbref_copy_1_1)
DECL_COPY(sbref_copy_2_2)
DECL_COPY(sbref_copy_3_3)
DECL_COPY(sbref_copy_4_4)
DECL_COPY(sbref_copy_0_1)
DECL_COPY(sbref_copy_1_0)
DECL_COPY(sbref_copy_2_3)
DECL_COPY(sbref_copy_3_2)
#    define DECL_CONTR(N)                                                                          \
        int N(const double *, int, const int *, const int *, const int *, const int *, int,       \
              const char *, int, const void **, int, const int *, const int *, const int *,       \
              const int *, int, const char *, int, const void **, const double *, int,            \
              const int *, const int *, const int *, const int *, int, const char *, void **,     \
              int);
DECL_CONTR(sbref_contraction_0)
DECL_CONTR(sbref_contraction_1)
DECL_CONTR(sbref_contraction_2)
DECL_CONTR(sbref_contraction_3)

/// mask0/mask1: NULL, or one MaskType (float) array per component, laid out like the component
int sbref_copy_masked(int dtype0, int dtype1, const double *alpha, int nd0, const int *p0,
                      int ncomp0, const char *o0, const int *from0, const int *size0,
                      const int *dim0, const void **v0, const float **mask0, int nd1,
                      const int *p1, int ncomp1, const char *o1, const int *from1, const int *dim1,
                      void **v1, const float **mask1, int co, int copyadd) {
#    define CASE(A, B)                                                                             \
        if (dtype0 == A && dtype1 == B)                                                            \
            return sbref_copy_##A##_##B(alpha, nd0, p0, ncomp0, o0, from0, size0, dim0, v0, nd1,   \
                                        p1, ncomp1, o1, from1, dim1, v1, co, copyadd, mask0,       \
                                        mask1);
    CASE(0, 0) CASE(1, 1) CASE(2, 2) CASE(3, 3) CASE(4, 4) CASE(0, 1) CASE(1, 0) CASE(2, 3)
        CASE(3, 2)
#    undef CASE
    g_err = "sbref_copy: unsupported type combination";
    return 1;
}

int sbref_copy(int dtype0, int dtype1, const double *alpha, int nd0, const int *p0, int ncomp0,
               const char *o0, const int *from0, const int *size0, const int *dim0,
               const void **v0, int nd1, const int *p1, int ncomp1, const char *o1,
               const int *from1, const int *dim1, void **v1, int co, int copyadd) {
    return sbref_copy_masked(dtype0, dtype1, alpha, nd0, p0, ncomp0, o0, from0, size0, dim0, v0,
                             nullptr, nd1, p1, ncomp1, o1, from1, dim1, v1, nullptr, co, copyadd);
}

int sbref_contraction(int dtype, const double *alpha, int nd0, const int *p0, const int *from0,
                      const int *size0, const int *dim0, int ncomp0, const char *o0, int conj0,
                      const void **v0, int nd1, const int *p1, const int *from1, const int *size1,
                      const int *dim1, int ncomp1, const char *o1, int conj1, const void **v1,
                      const double *beta, int ndo, const int *pr, const int *fromr,
                      const int *sizer, const int *dimr, int ncompr, const char *o_r, void **vr,
                      int co) {
#    define CASE(A)                                                                                \
        if (dtype == A)                                                                            \
            return sbref_contraction_##A(alpha, nd0, p0, from0, size0, dim0, ncomp0, o0, conj0,    \
                                         v0, nd1, p1, from1, size1, dim1, ncomp1, o1, conj1, v1,   \
                                         beta, ndo, pr, fromr, sizer, dimr, ncompr, o_r, vr, co);
    CASE(0) CASE(1) CASE(2) CASE(3)
#    undef CASE
    g_err = "sbref_contraction: unsupported type";
    return 1;
}

#    define NDSWITCH(F, ...)                                                                       \
        switch (nd) {                                                                              \
        case 1: F<1>(__VA_ARGS__); break;                                                          \
        case 2: F<2>(__VA_ARGS__); break;                                                          \
        case 3: F<3>(__VA_ARGS__); break;                                                          \
        case 4: F<4>(__VA_ARGS__); break;                                                          \
        case 5: F<5>(__VA_ARGS__); break;                                                          \
        case 6: F<6>(__VA_ARGS__); break;                                                          \
        case 7: F<7>(__VA_ARGS__); break;                                                          \
        case 8: F<8>(__VA_ARGS__); break;                                                          \
        default: throw std::runtime_error("sbref: unsupported number of dimensions");             \
        }

/// out: [nprocs*ncomponents][2][nd]
int sbref_basic_partitioning(int nd, const char *order, const int *dim, const int *procs,
                             const char *dist_labels, int nprocs, int ncomponents, int *out) {
    SBREF_TRY(NDSWITCH(bp, order, dim, procs, dist_labels, nprocs, ncomponents, out));
}

/// out: [nprocs][2][nd]
int sbref_basic_partitioning_ext(int nd, const int *dim, const int *procs, int nprocs,
                                 int replicate, const int *ext_power, int *out) {
    SBREF_TRY(NDSWITCH(bpe, dim, procs, nprocs, replicate, ext_power, out));
}

int sbref_partitioning_distributed_procs(int nd, const char *order, const int *dim,
                                         const char *dist_labels, int nprocs, int *out) {
    SBREF_TRY(NDSWITCH(pdp, order, dim, dist_labels, nprocs, out));
}

/// out must hold up to 2*nd... entries [n][2][nd]; the caller passes room for 3^nd boxes
int sbref_make_hole(int nd, const int *from, const int *size, const int *hole_from,
                    const int *hole_size, const int *dim, int *out, int *nout) {
    SBREF_TRY(NDSWITCH(mh, from, size, hole_from, hole_size, dim, out, nout));
}

int sbref_clear_caches() { SBREF_TRY(clearCaches()); }
#endif // SBREF_PART_MAIN
}
