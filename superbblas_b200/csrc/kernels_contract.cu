// Contraction kernels for sm_100a:  vr = alpha * sum_K f0(v0) * f1(v1) + beta * vr
//
// Replaces the reference's lowering "permute both operands -> xgemm_batch_strided -> add-copy"
// (tensor.h:1271-1429 suggested_orders_for_contraction, tensor.h:1475-1598
// local_contraction_normalized, blas.h:662-810 cublasGemmStridedBatchedEx).  Here the operands are
// never re-laid-out: every label keeps its own stride (sbk_contract_desc) and the tile loaders
// gather straight from the caller's layout, so the permutation is fused into the GEMM.
//
// Kernels:
//  * contract_mma_kernel (double, complex double; float and complex float with their tiles kept in
//    float and widened at the fragment loads): FP64 tensor-core path.  tcgen05 has no f64 kind,
//    so on Blackwell FP64 matrix math is issued as warp-level DMMA (mma.sync.m8n8k4.f64).  A CTA
//    owns a 64x64 output tile of one batch entry and one K-slice; operands flow
//    global --cp.async (16 B per complex element, any stride)--> 4-stage shared-memory ring -->
//    LDS.128 fragments --> DMMA, a complex product being four real DMMAs (the sign of the
//    imaginary parts carries conj and the minus of i*i).  K is split across CTAs so that the grid
//    fills 148 SMs x 2 CTAs for several waves even when the output is a single tile per batch
//    entry (distillation shapes: M=N=64..128, K~1e5); partial tiles go to a workspace and
//  * contract_reduce_kernel sums them in a fixed order (deterministic), applies alpha/beta/output
//    strides (the reference's final add-copy, dist.h:3184, fused here).
//  * contract_simt_kernel: any type, any shape; one thread per output element, K loop in registers,
//    float types accumulate in double.  Used for short contractions (e.g. site-wise colour-spin
//    contractions with M=N=1) and small tiles with many outputs.
//  * contract_row_kernel / contract_row_smem_kernel (contract_row.hpp: short K and one small free
//    group; the small operand staged in shared memory when a CTA's rows share it) and
//    contract_dot_*_kernel (contract_dot.hpp: long K and two small free groups; CTA tree + a warp per
//    output in the second pass): the skinny shapes the reference hands to GEMV / dot calls
//    (blas.h:686-699).  Bodies also run on the CPU (tests/test_row_kernel_emulation.py).
//  * complex float with a long contiguous K goes to the tcgen05 kernel (kernels_contract_tc.cu).
#include "contract_dot.hpp"
#include "contract_row.hpp"
#include "contract_tc.hpp"
#include "kernels.hpp"
#include "runtime.hpp"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

namespace sbb {

    namespace {

        constexpr int GD = SBK_MAX_GROUP_DIMS;

        struct Group {
            int n;
            int size[GD];
            long long s0[GD], s1[GD], sr[GD];
            long long vol;
        };

        struct ContractParams {
            Group T, M, N, K;
            int conj0, conj1;
            int bn; // columns of an output tile of the mma kernel
            int out_order[3]; // generic kernel: groups (0 = T, 1 = M, 2 = N) from fastest to slowest thread index
            // mma kernel
            int mtiles, ntiles, ksplit, ksteps; // ksteps = ceil(K/BK)
            int a_sr, a_sk, b_sr, b_sk;         // shared-memory strides (elements) of the operand tiles
            int a_kfast, b_kfast;               // loader enumeration: 1 = k fastest, 0 = row fastest
        };

        __host__ __device__ inline long long group_offset(const Group &g, long long idx,
                                                          const long long *stride) {
            long long off = 0;
#pragma unroll
            for (int d = 0; d < GD; ++d)
                if (d < g.n) {
                    const long long c = idx % g.size[d];
                    idx /= g.size[d];
                    off += c * stride[d];
                }
            return off;
        }

        // ---- complex helpers ---------------------------------------------------------------------

        template <typename T> struct Acc { using type = T; };
        template <> struct Acc<float> { using type = double; };
        template <> struct Acc<float2> { using type = double2; };

        __device__ __forceinline__ double widen(float x) { return (double)x; }
        __device__ __forceinline__ double widen(double x) { return x; }
        __device__ __forceinline__ double2 widen(float2 x) { return make_double2(x.x, x.y); }
        __device__ __forceinline__ double2 widen(double2 x) { return x; }
        __device__ __forceinline__ double cj(double x) { return x; }
        __device__ __forceinline__ double2 cj(double2 x) { return make_double2(x.x, -x.y); }
        __device__ __forceinline__ void fma_acc(double &acc, double a, double b) { acc = fma(a, b, acc); }
        __device__ __forceinline__ void fma_acc(double2 &acc, double2 a, double2 b) {
            acc.x = fma(a.x, b.x, acc.x);
            acc.x = fma(-a.y, b.y, acc.x);
            acc.y = fma(a.x, b.y, acc.y);
            acc.y = fma(a.y, b.x, acc.y);
        }
        __device__ __forceinline__ double mulc(double a, double b) { return a * b; }
        __device__ __forceinline__ double2 mulc(double2 a, double2 b) {
            return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
        }
        __device__ __forceinline__ double addc(double a, double b) { return a + b; }
        __device__ __forceinline__ double2 addc(double2 a, double2 b) {
            return make_double2(a.x + b.x, a.y + b.y);
        }
        template <typename T> __device__ __forceinline__ T narrow(typename Acc<T>::type x);
        template <> __device__ __forceinline__ float narrow<float>(double x) { return (float)x; }
        template <> __device__ __forceinline__ double narrow<double>(double x) { return x; }
        template <> __device__ __forceinline__ float2 narrow<float2>(double2 x) {
            return make_float2((float)x.x, (float)x.y);
        }
        template <> __device__ __forceinline__ double2 narrow<double2>(double2 x) { return x; }
        template <typename A> __device__ __forceinline__ A zero_acc();
        template <> __device__ __forceinline__ double zero_acc<double>() { return 0.0; }
        template <> __device__ __forceinline__ double2 zero_acc<double2>() {
            return make_double2(0.0, 0.0);
        }
        template <typename A> __device__ __forceinline__ bool is_zero_acc(A x);
        template <> __device__ __forceinline__ bool is_zero_acc<double>(double x) { return x == 0; }
        template <> __device__ __forceinline__ bool is_zero_acc<double2>(double2 x) {
            return x.x == 0 && x.y == 0;
        }

        // ---- generic kernel ------------------------------------------------------------------------

        template <typename T>
        __global__ void __launch_bounds__(256)
            contract_simt_kernel(const __grid_constant__ ContractParams p, const T *__restrict__ v0,
                                 const T *__restrict__ v1, T *vr, typename Acc<T>::type alpha,
                                 typename Acc<T>::type beta) {
            using A = typename Acc<T>::type;
            const long long total = p.T.vol * p.M.vol * p.N.vol;
            for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
                 idx += (long long)gridDim.x * blockDim.x) {
                // output enumeration: the group with the smallest output stride fastest, so that the
                // stores are coalesced (default order: n fastest, then m, then t)
                long long n = idx % p.N.vol, m = (idx / p.N.vol) % p.M.vol,
                          t = idx / (p.N.vol * p.M.vol);
                if (p.out_order[0] != 2 || p.out_order[1] != 1)
                    rowk::output_index(p.out_order, p.T.vol, p.M.vol, p.N.vol, idx, t, m, n);
                const long long o0 = group_offset(p.T, t, p.T.s0) + group_offset(p.M, m, p.M.s0);
                const long long o1 = group_offset(p.T, t, p.T.s1) + group_offset(p.N, n, p.N.s1);
                const long long orr = group_offset(p.T, t, p.T.sr) + group_offset(p.M, m, p.M.sr) +
                                      group_offset(p.N, n, p.N.sr);
                A acc = zero_acc<A>();
                if (p.K.n <= 1) {
                    const long long k0 = p.K.n ? p.K.s0[0] : 0, k1 = p.K.n ? p.K.s1[0] : 0;
                    for (long long k = 0; k < p.K.vol; ++k) {
                        A a = widen(v0[o0 + k * k0]), b = widen(v1[o1 + k * k1]);
                        if (p.conj0) a = cj(a);
                        if (p.conj1) b = cj(b);
                        fma_acc(acc, a, b);
                    }
                } else {
                    for (long long k = 0; k < p.K.vol; ++k) {
                        A a = widen(v0[o0 + group_offset(p.K, k, p.K.s0)]),
                          b = widen(v1[o1 + group_offset(p.K, k, p.K.s1)]);
                        if (p.conj0) a = cj(a);
                        if (p.conj1) b = cj(b);
                        fma_acc(acc, a, b);
                    }
                }
                A r = mulc(alpha, acc);
                if (!is_zero_acc(beta)) r = addc(r, mulc(beta, widen(vr[orr])));
                vr[orr] = narrow<T>(r);
            }
        }

        // ---- row kernel (short contraction, one small free group; body in contract_row.hpp) ----------

        template <typename T>
        __global__ void __launch_bounds__(128)
            contract_row_kernel(const __grid_constant__ rowk::RowParams p, const T *__restrict__ va,
                                const T *__restrict__ vb, T *vr, typename rowk::Acc<T>::type alpha,
                                typename rowk::Acc<T>::type beta) {
            for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < p.rows;
                 row += (long long)gridDim.x * blockDim.x)
                rowk::row_body<T>(p, row, va, vb, vr, alpha, beta);
        }

        /// The same with the small operand staged in shared memory: when the 128 rows of a CTA share
        /// one K x S block of it (`inner` = rows per block is a multiple of 128), the block is loaded
        /// once per CTA and every product reads it with a broadcast LDS instead of a load through L1
        /// (S x K loads per row: they, not the big operand, bound the kernel -- 49152 x 16 x 16
        /// "update" shapes ran at 14 % of the HBM floor).
        template <typename T>
        __global__ void __launch_bounds__(128)
            contract_row_smem_kernel(const __grid_constant__ rowk::RowParams p, const T *__restrict__ va,
                                     const T *__restrict__ vb, T *vr, typename rowk::Acc<T>::type alpha,
                                     typename rowk::Acc<T>::type beta) {
            using A = typename rowk::Acc<T>::type;
            using F = typename rowk::Fast<T>::type;
            __shared__ F sb[rowk::KMAX * rowk::SMAX];
            const long long nblocks = (p.rows + 127) / 128;
            for (long long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
                const long long row = blk * 128 + threadIdx.x;
                // offsets of this row (the small operand's offset is the same for the whole CTA)
                long long oa = 0, ob = 0, orr = 0, rem = row < p.rows ? row : p.rows - 1;
#pragma unroll 1
                for (int d = 0; d < p.nd; ++d) {
                    const long long c = rem % p.size[d];
                    rem /= p.size[d];
                    oa += c * p.sa[d], ob += c * p.sb[d], orr += c * p.sr[d];
                }
                __syncthreads(); // the previous block's products are done with sb
                for (int i = threadIdx.x; i < p.nk * p.ns; i += 128) {
                    F b = vb[ob + p.koff_b[i / p.ns] + p.soff_b[i % p.ns]];
                    if (p.conj_b) b = rowk::cj(b);
                    sb[i] = b;
                }
                __syncthreads();
                if (row >= p.rows) continue;
                F acc[rowk::SMAX];
#pragma unroll
                for (int s = 0; s < rowk::SMAX; ++s) rowk::set_zero(acc[s]);
                constexpr int KB = 8;
                for (int k0 = 0; k0 < p.nk; k0 += KB) {
                    F a[KB];
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                        if (k0 + j < p.nk) {
                            a[j] = va[oa + p.koff_a[k0 + j]];
                            if (p.conj_a) a[j] = rowk::cj(a[j]);
                        }
#pragma unroll
                    for (int j = 0; j < KB; ++j)
                        if (k0 + j < p.nk) {
                            const F *bk = sb + (k0 + j) * p.ns;
#pragma unroll
                            for (int s = 0; s < rowk::SMAX; ++s)
                                if (s < p.ns) rowk::fma_acc(acc[s], a[j], bk[s]);
                        }
                }
#pragma unroll
                for (int s = 0; s < rowk::SMAX; ++s)
                    if (s < p.ns) {
                        A r = rowk::mulc(alpha, rowk::widen(acc[s]));
                        T *w = vr + orr + p.soff_r[s];
                        if (!rowk::is_zero(beta)) r = rowk::addc(r, rowk::mulc(beta, rowk::widen(*w)));
                        rowk::narrow(r, *w);
                    }
            }
        }

        // ---- dot kernel (long contraction, both free groups small; bodies in contract_dot.hpp) -------

        __device__ __forceinline__ double shfl_down(double v, int off) { return __shfl_down_sync(0xffffffffu, v, off); }
        __device__ __forceinline__ double2 shfl_down(double2 v, int off) {
            return make_double2(__shfl_down_sync(0xffffffffu, v.x, off), __shfl_down_sync(0xffffffffu, v.y, off));
        }

        template <typename T>
        __global__ void __launch_bounds__(dotk::CTA)
            contract_dot_partial_kernel(const __grid_constant__ dotk::DotParams p,
                                        const T *__restrict__ va, const T *__restrict__ vb,
                                        typename rowk::Acc<T>::type *__restrict__ ws) {
            using A = typename rowk::Acc<T>::type;
            constexpr int SB = dotk::SB;
            const long long thread = blockIdx.x * (long long)blockDim.x + threadIdx.x;
            if (!dotk::cta_tree(p)) { // few slices: every thread writes its own block
                if (thread < dotk::threads_of(p)) dotk::dot_partial<T>(p, thread, va, vb, ws);
                return;
            }
            // the CTA holds 128 slices of the same outputs: add them here (fixed tree: deterministic)
            typename rowk::Fast<T>::type acc[SB][SB];
            dotk::dot_partial_acc<T>(p, thread, va, vb, acc);
            __shared__ A part[dotk::CTA / 32][SB * SB];
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            // (only the live entries of the block: for m = n = 1 the 15 dead ones cost 300 shuffles
            // per thread, most of the kernel; the condition is uniform over the CTA)
            const int pgi = (int)((blockIdx.x / (p.slices / dotk::CTA)) % (p.pgm * p.pgn));
            const int mlive = min(SB, p.m - (pgi / p.pgn) * SB), nlive = min(SB, p.n - (pgi % p.pgn) * SB);
#pragma unroll
            for (int q = 0; q < SB * SB; ++q) {
                A v = rowk::widen(acc[q / SB][q % SB]);
                if (q / SB < mlive && q % SB < nlive) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v = rowk::addc(v, shfl_down(v, off));
                }
                if (lane == 0) part[warp][q] = v;
            }
            __syncthreads();
            if (threadIdx.x < SB * SB) {
                A v = part[0][threadIdx.x];
#pragma unroll
                for (int w = 1; w < dotk::CTA / 32; ++w) v = rowk::addc(v, part[w][threadIdx.x]);
                ws[blockIdx.x * (long long)(SB * SB) + threadIdx.x] = v;
            }
        }

        /// Second pass of the dot kernel: one WARP per output -- the lanes stride over the slices left
        /// by the first pass and add their sums with a fixed shuffle tree (deterministic).  (One thread
        /// per output walked the slices one dependent load after the other: 38 of the 40 us of an
        /// m = n = 1, k = 49152, batch 32 call.)
        template <typename T>
        __global__ void __launch_bounds__(256)
            contract_dot_reduce_kernel(const __grid_constant__ dotk::DotParams p,
                                       const typename rowk::Acc<T>::type *__restrict__ ws, T *vr,
                                       typename rowk::Acc<T>::type alpha,
                                       typename rowk::Acc<T>::type beta) {
            using A = typename rowk::Acc<T>::type;
            constexpr int SB = dotk::SB;
            const long long total = dotk::outputs_of(p);
            const int lane = threadIdx.x & 31, nsl = dotk::ws_slices(p);
            const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
            for (long long out = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; out < total;
                 out += nwarps) {
                const int nn = (int)(out % p.n);
                long long rem = out / p.n;
                const int mm = (int)(rem % p.m);
                long long t = rem / p.m;
                const int pg = (mm / SB) * p.pgn + nn / SB;
                const A *src = ws + ((t * (p.pgm * p.pgn) + pg) * (long long)nsl) * (SB * SB) +
                               (mm % SB) * SB + nn % SB;
                A acc;
                rowk::set_zero(acc);
                for (int s = lane; s < nsl; s += 32) acc = rowk::addc(acc, src[(long long)s * (SB * SB)]);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc = rowk::addc(acc, shfl_down(acc, off));
                if (lane == 0) {
                    long long orr = p.moff_r[mm] + p.noff_r[nn];
                    for (int d = 0; d < p.nd_t; ++d) {
                        const long long c = t % p.size_t_[d];
                        t /= p.size_t_[d];
                        orr += c * p.sr_t[d];
                    }
                    A r = rowk::mulc(alpha, acc);
                    T *w = vr + orr;
                    if (!rowk::is_zero(beta)) r = rowk::addc(r, rowk::mulc(beta, rowk::widen(*w)));
                    rowk::narrow(r, *w);
                }
            }
        }

        // ---- FP64 tensor-core kernel -----------------------------------------------------------------

        constexpr int BM = 64, MMA_THREADS = 128;
        // BN (tile columns), pipeline depth and resident CTAs are template parameters: 64/4/2 and 32/4/3 are built

        __device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(d0), "+d"(d1)
                         : "d"(a), "d"(b));
        }

        template <int BYTES>
        __device__ __forceinline__ void cp_async(void *smem, const void *gmem, bool valid) {
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(smem);
            const int src = valid ? BYTES : 0; // src-size 0: the destination is zero filled
            asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(saddr), "l"(gmem),
                         "n"(BYTES), "r"(src));
        }
        __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
        template <int N> __device__ __forceinline__ void cp_async_wait() {
            asm volatile("cp.async.wait_group %0;" ::"n"(N));
        }

        template <typename T> struct Frag; // fragment element as loaded from shared memory
        template <> struct Frag<double> {
            static constexpr bool cplx = false;
        };
        template <> struct Frag<double2> {
            static constexpr bool cplx = true;
        };
        // float and complex float operands stay in their own type up to shared memory (half the HBM
        // and shared-memory traffic) and are widened when the fragments are read: the products and
        // the whole K sum are done by the FP64 tensor pipe, only the final result is rounded to float
        template <> struct Frag<float> {
            static constexpr bool cplx = false;
        };
        template <> struct Frag<float2> {
            static constexpr bool cplx = true;
        };

        /// Offset of contracted index k in an operand
        __device__ __forceinline__ long long k_offset(const Group &K, long long k,
                                                      const long long *stride) {
            if (K.n == 1) return k * stride[0];
            return group_offset(K, k, stride);
        }

        template <typename T, int BN, int BK, int STAGES, int MINB>
        __global__ void __launch_bounds__(MMA_THREADS, MINB)
            contract_mma_kernel(const __grid_constant__ ContractParams p, const T *__restrict__ v0,
                                const T *__restrict__ v1, typename Acc<T>::type *__restrict__ ws) {
            using A = typename Acc<T>::type; // double or double2: what the partial tiles are stored as
            constexpr bool CPLX = Frag<T>::cplx;
            constexpr int ACC = CPLX ? 4 : 2; // doubles per 8x8 block per lane
            constexpr int WN = BN / 16;       // 8-column blocks per warp (warp tile = 32 x BN/2)
            extern __shared__ __align__(16) unsigned char smem_raw[];
            // stage layout: [A tile | B tile], sizes fixed by the host (a_stage, b_stage elements)
            const int a_stage = p.a_kfast ? BM * p.a_sr : BK * p.a_sk;
            const int b_stage = p.b_kfast ? BN * p.b_sr : BK * p.b_sk;
            T *smem = reinterpret_cast<T *>(smem_raw);
            const int stage_elems = a_stage + b_stage;

            const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
            const int wm = warp >> 1, wn = warp & 1; // 2x2 warps, 32 x BN/2 each

            // ---- which tile / K slice ------------------------------------------------------------
            long long bid = blockIdx.x;
            const int nt = (int)(bid % p.ntiles);
            bid /= p.ntiles;
            const int mt = (int)(bid % p.mtiles);
            bid /= p.mtiles;
            const int ks = (int)(bid % p.ksplit);
            const long long t = bid / p.ksplit;
            const int kstep0 = (int)((long long)p.ksteps * ks / p.ksplit);
            const int kstep1 = (int)((long long)p.ksteps * (ks + 1) / p.ksplit);
            const int nsteps = kstep1 - kstep0;

            const T *a_base = v0 + group_offset(p.T, t, p.T.s0);
            const T *b_base = v1 + group_offset(p.T, t, p.T.s1);

            // ---- loader slots: 4 elements of A and 4 of B per thread and stage ----------------------
            constexpr int A_SLOTS = BM * BK / MMA_THREADS, B_SLOTS = BN * BK / MMA_THREADS;
            long long a_row[A_SLOTS], b_row[B_SLOTS];
            int a_kk[A_SLOTS], b_kk[B_SLOTS], a_sm[A_SLOTS], b_sm[B_SLOTS];
#pragma unroll
            for (int i = 0; i < A_SLOTS; ++i) {
                const int slot = tid + i * MMA_THREADS;
                const int r = p.a_kfast ? slot / BK : slot % BM, kk = p.a_kfast ? slot % BK : slot / BM;
                const long long m = min((long long)mt * BM + r, p.M.vol - 1); // clamp: rows past M are never read back
                a_row[i] = group_offset(p.M, m, p.M.s0);
                a_kk[i] = kk;
                a_sm[i] = r * p.a_sr + kk * p.a_sk;
            }
#pragma unroll
            for (int i = 0; i < B_SLOTS; ++i) {
                const int slot = tid + i * MMA_THREADS;
                const int r = p.b_kfast ? slot / BK : slot % BN, kk = p.b_kfast ? slot % BK : slot / BN;
                const long long n = min((long long)nt * BN + r, p.N.vol - 1);
                b_row[i] = group_offset(p.N, n, p.N.s1);
                b_kk[i] = kk;
                b_sm[i] = a_stage + r * p.b_sr + kk * p.b_sk;
            }

            // When the contracted labels merge into one dimension (the common case) every slot walks
            // its row with a constant pointer increment; otherwise the K offset is decomposed per load.
            const bool k_linear = p.K.n <= 1;
            const T *a_ptr[A_SLOTS], *b_ptr[B_SLOTS];
            const long long a_step = (p.K.n ? p.K.s0[0] : 0) * BK, b_step = (p.K.n ? p.K.s1[0] : 0) * BK;
#pragma unroll
            for (int i = 0; i < A_SLOTS; ++i)
                a_ptr[i] = a_base + a_row[i] + ((long long)kstep0 * BK + a_kk[i]) * (p.K.n ? p.K.s0[0] : 0);
#pragma unroll
            for (int i = 0; i < B_SLOTS; ++i)
                b_ptr[i] = b_base + b_row[i] + ((long long)kstep0 * BK + b_kk[i]) * (p.K.n ? p.K.s1[0] : 0);
            auto load_stage = [&](int stage, int kstep) {
                T *s = smem + (size_t)stage * stage_elems;
                const long long kbase = (long long)kstep * BK;
                if (k_linear) {
                    // stages are loaded in k order, so the pointers simply advance
                    const bool tail = kbase + BK > p.K.vol;
#pragma unroll
                    for (int i = 0; i < A_SLOTS; ++i) {
                        const bool ok = !tail || kbase + a_kk[i] < p.K.vol;
                        cp_async<sizeof(T)>(s + a_sm[i], ok ? a_ptr[i] : a_base, ok);
                        a_ptr[i] += a_step;
                    }
#pragma unroll
                    for (int i = 0; i < B_SLOTS; ++i) {
                        const bool ok = !tail || kbase + b_kk[i] < p.K.vol;
                        cp_async<sizeof(T)>(s + b_sm[i], ok ? b_ptr[i] : b_base, ok);
                        b_ptr[i] += b_step;
                    }
                    return;
                }
#pragma unroll
                for (int i = 0; i < A_SLOTS; ++i) {
                    const long long k = kbase + a_kk[i];
                    const bool ok = k < p.K.vol;
                    const long long off = ok ? a_row[i] + k_offset(p.K, k, p.K.s0) : 0;
                    cp_async<sizeof(T)>(s + a_sm[i], a_base + off, ok);
                }
#pragma unroll
                for (int i = 0; i < B_SLOTS; ++i) {
                    const long long k = kbase + b_kk[i];
                    const bool ok = k < p.K.vol;
                    const long long off = ok ? b_row[i] + k_offset(p.K, k, p.K.s1) : 0;
                    cp_async<sizeof(T)>(s + b_sm[i], b_base + off, ok);
                }
            };

            // ---- accumulators: 4x4 blocks of 8x8 per warp ---------------------------------------------
            double acc[4][WN][ACC];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < WN; ++j)
#pragma unroll
                    for (int c = 0; c < ACC; ++c) acc[i][j][c] = 0.0;

            // fragment addresses inside a stage: element (row = base + lane/4, k = lane%4)
            int a_frag[4], b_frag[WN];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                a_frag[i] = (wm * 32 + i * 8 + (lane >> 2)) * p.a_sr + (lane & 3) * p.a_sk;
#pragma unroll
            for (int j = 0; j < WN; ++j)
                b_frag[j] = a_stage + (wn * (BN / 2) + j * 8 + (lane >> 2)) * p.b_sr + (lane & 3) * p.b_sk;
            // conj and the minus of i*i are sign flips of the imaginary fragments.  They are done on
            // the integer pipe (XOR of the sign bit): as FP64 multiplies / negations (8 per 64 DMMAs)
            // they shared the FP64 pipe with the DMMAs and drew 23 % of the stall samples
            // (profiles/r2_contract_mma_ncu.txt: DADD), with the tensor pipe at 88 %.
            const unsigned sa = p.conj0 ? 0x80000000u : 0u, sb = p.conj1 ? 0x80000000u : 0u;
            auto flip = [](double x, unsigned m) {
                return __hiloint2double(__double2hiint(x) ^ (int)m, __double2loint(x));
            };

            // ---- pipeline -----------------------------------------------------------------------------
#pragma unroll
            for (int s = 0; s < STAGES - 1; ++s) {
                if (s < nsteps) load_stage(s, kstep0 + s);
                cp_async_commit();
            }
            for (int it = 0; it < nsteps; ++it) {
                cp_async_wait<STAGES - 2>();
                __syncthreads();
                const T *s = smem + (size_t)(it % STAGES) * stage_elems;
#pragma unroll
                for (int k4 = 0; k4 < BK / 4; ++k4) {
                    if (k4 == (BK / 4 > 1 ? 1 : 0)) {
                        // the next stage's loads are issued in the shadow of the first block of DMMAs
                        // (the stage they overwrite was released by the barrier above)
                        const int nxt = it + STAGES - 1;
                        if (nxt < nsteps) load_stage(nxt % STAGES, kstep0 + nxt);
                        cp_async_commit();
                    }
                    if constexpr (CPLX) {
                        double ar[4], ai[4], nai[4], br[WN], bi[WN];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const double2 a = widen(s[a_frag[i] + k4 * 4 * p.a_sk]);
                            ar[i] = a.x, ai[i] = flip(a.y, sa), nai[i] = flip(a.y, sa ^ 0x80000000u);
                        }
#pragma unroll
                        for (int j = 0; j < WN; ++j) {
                            const double2 b = widen(s[b_frag[j] + k4 * 4 * p.b_sk]);
                            br[j] = b.x, bi[j] = flip(b.y, sb);
                        }
                        // four passes over the 16 blocks: the two DMMAs that accumulate into the
                        // same registers are 32 instructions apart, so none waits for the other
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < WN; ++j) {
                                dmma(acc[i][j][0], acc[i][j][1], ar[i], br[j]);
                                dmma(acc[i][j][2], acc[i][j][3], ar[i], bi[j]);
                            }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < WN; ++j) {
                                dmma(acc[i][j][0], acc[i][j][1], nai[i], bi[j]);
                                dmma(acc[i][j][2], acc[i][j][3], ai[i], br[j]);
                            }
                    } else {
                        double a[4], b[WN];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            a[i] = widen(s[a_frag[i] + k4 * 4 * p.a_sk]);
#pragma unroll
                        for (int j = 0; j < WN; ++j)
                            b[j] = widen(s[b_frag[j] + k4 * 4 * p.b_sk]);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    }
                }
            }
            cp_async_wait<0>();

            // ---- partial tile to the workspace: ws[((t*mt*nt tile) * ksplit + ks)][m][n] -------------
            const long long tile_id = ((t * p.mtiles + mt) * p.ntiles + nt) * p.ksplit + ks;
            A *out = ws + tile_id * (long long)(BM * BN);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < WN; ++j) {
                    const int row = wm * 32 + i * 8 + (lane >> 2);
                    const int col = wn * (BN / 2) + j * 8 + (lane & 3) * 2;
                    if constexpr (CPLX) {
                        double4 v0 = make_double4(acc[i][j][0], acc[i][j][2], acc[i][j][1],
                                                  acc[i][j][3]);
                        *reinterpret_cast<double4 *>(out + row * BN + col) = v0;
                    } else {
                        *reinterpret_cast<double2 *>(out + row * BN + col) =
                            make_double2(acc[i][j][0], acc[i][j][1]);
                    }
                }
        }

        /// Sum the K-slices of every output tile in a fixed order and write alpha*sum + beta*vr
        template <typename T>
        __global__ void __launch_bounds__(256)
            contract_reduce_kernel(const __grid_constant__ ContractParams p,
                                   const typename Acc<T>::type *__restrict__ ws, T *vr,
                                   typename Acc<T>::type alpha, typename Acc<T>::type beta) {
            using A = typename Acc<T>::type;
            const long long total = p.T.vol * p.M.vol * p.N.vol;
            for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
                 idx += (long long)gridDim.x * blockDim.x) {
                const long long n = idx % p.N.vol, m = (idx / p.N.vol) % p.M.vol,
                                t = idx / (p.N.vol * p.M.vol);
                const int BN = p.bn;
                const int mt = (int)(m / BM), nt = (int)(n / BN);
                const long long tile0 = ((t * p.mtiles + mt) * p.ntiles + nt) * p.ksplit;
                const A *src = ws + tile0 * (long long)(BM * BN) + (m % BM) * BN + (n % BN);
                A acc = src[0];
                for (int s = 1; s < p.ksplit; ++s) acc = addc(acc, src[(long long)s * BM * BN]);
                const long long orr = group_offset(p.T, t, p.T.sr) + group_offset(p.M, m, p.M.sr) +
                                      group_offset(p.N, n, p.N.sr);
                A r = mulc(alpha, acc);
                if (!is_zero_acc(beta)) r = addc(r, mulc(beta, widen(vr[orr])));
                vr[orr] = narrow<T>(r);
            }
        }

        // ---- host ----------------------------------------------------------------------------------

        /// Copy one label group, dropping extent-1 labels and merging neighbours that stay affine in
        /// every tensor; dims are sorted by the stride in `key` (0: s0, 1: s1, 2: sr)
        Group make_group(const sbk_contract_dim *d, int n, int key) {
            std::vector<sbk_contract_dim> v;
            for (int i = 0; i < n; ++i) {
                if (d[i].size <= 0) {
                    Group g;
                    std::memset(&g, 0, sizeof g);
                    g.vol = 0;
                    return g;
                }
                if (d[i].size > 1) v.push_back(d[i]);
            }
            auto keyof = [&](const sbk_contract_dim &x) { return key == 0 ? x.s0 : key == 1 ? x.s1 : x.sr; };
            std::stable_sort(v.begin(), v.end(), [&](const sbk_contract_dim &a, const sbk_contract_dim &b) {
                return keyof(a) < keyof(b);
            });
            std::vector<sbk_contract_dim> m;
            for (const auto &x : v) {
                if (!m.empty()) {
                    auto &l = m.back();
                    const long long s = l.size;
                    if (l.s0 * s == x.s0 && l.s1 * s == x.s1 && l.sr * s == x.sr &&
                        s * (long long)x.size < (1ll << 31)) {
                        l.size *= x.size;
                        continue;
                    }
                }
                m.push_back(x);
            }
            if ((int)m.size() > GD) throw std::runtime_error("contraction: too many labels in a group");
            Group g;
            std::memset(&g, 0, sizeof g);
            g.n = (int)m.size();
            g.vol = 1;
            for (int i = 0; i < g.n; ++i) {
                g.size[i] = m[i].size, g.s0[i] = m[i].s0, g.s1[i] = m[i].s1, g.sr[i] = m[i].sr;
                g.vol *= m[i].size;
            }
            return g;
        }

        template <typename T> T scalar_of(const double *a);
        template <> double scalar_of<double>(const double *a) { return a[0]; }
        template <> double2 scalar_of<double2>(const double *a) { return make_double2(a[0], a[1]); }

        template <typename T>
        void launch_row(const rowk::RowParams &rp, bool swapped, const double *alpha, const void *v0,
                        const void *v1, const double *beta, void *vr, int device, cudaStream_t stream) {
            using A = typename Acc<T>::type;
            const unsigned grid =
                (unsigned)std::min<long long>((rp.rows + 127) / 128, (long long)sm_count(device) * 12);
            // rows that share one block of the small operand: the leading row dims it does not depend on
            long long inner = 1;
            for (int d = 0; d < rp.nd && rp.sb[d] == 0; ++d) inner *= rp.size[d];
            if (inner % 128 == 0 && rp.ns * rp.nk >= 4)
                contract_row_smem_kernel<T><<<grid, 128, 0, stream>>>(rp, (const T *)(swapped ? v1 : v0),
                                                                     (const T *)(swapped ? v0 : v1), (T *)vr,
                                                                     scalar_of<A>(alpha), scalar_of<A>(beta));
            else
                contract_row_kernel<T><<<grid, 128, 0, stream>>>(rp, (const T *)(swapped ? v1 : v0),
                                                                (const T *)(swapped ? v0 : v1), (T *)vr,
                                                                scalar_of<A>(alpha), scalar_of<A>(beta));
            count_launch();
            cuda_check(cudaGetLastError(), "contract_row_kernel launch");
        }

        template <typename T>
        void launch_dot(const dotk::DotParams &dp, const double *alpha, const void *v0, const void *v1,
                        const double *beta, void *vr, int device, cudaStream_t stream) {
            using A = typename Acc<T>::type;
            const long long threads = dotk::threads_of(dp);
            A *ws = (A *)pool_alloc(device, (size_t)threads * dotk::SB * dotk::SB * sizeof(A));
            contract_dot_partial_kernel<T><<<(unsigned)((threads + dotk::CTA - 1) / dotk::CTA), dotk::CTA, 0, stream>>>(
                dp, (const T *)v0, (const T *)v1, ws);
            count_launch();
            cuda_check(cudaGetLastError(), "contract_dot_partial_kernel launch");
            const long long outs = dotk::outputs_of(dp); // one warp each
            const unsigned grid =
                (unsigned)std::min<long long>((outs * 32 + 255) / 256, (long long)sm_count(device) * 8);
            contract_dot_reduce_kernel<T><<<grid, 256, 0, stream>>>(dp, ws, (T *)vr, scalar_of<A>(alpha),
                                                                  scalar_of<A>(beta));
            count_launch();
            cuda_check(cudaGetLastError(), "contract_dot_reduce_kernel launch");
            pool_free(device, ws);
        }

        template <typename T>
        void launch_simt(const ContractParams &p, const double *alpha, const void *v0, const void *v1,
                         const double *beta, void *vr, int device, cudaStream_t stream) {
            using A = typename Acc<T>::type;
            const long long total = p.T.vol * p.M.vol * p.N.vol;
            const unsigned grid =
                (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count(device) * 16);
            contract_simt_kernel<T><<<grid, 256, 0, stream>>>(p, (const T *)v0, (const T *)v1, (T *)vr,
                                                             scalar_of<A>(alpha), scalar_of<A>(beta));
            count_launch();
            cuda_check(cudaGetLastError(), "contract_simt_kernel launch");
        }

        /// Shared-memory layout of an operand tile: k-contiguous operands are stored [row][k] with
        /// row stride = BK + pad, row-contiguous operands [k][row]; the pads make the fragment
        /// LDS conflict free (see the derivation in DESIGN.md).
        void tile_layout(bool kfast, int esize, int rows, int bk, int &sr, int &sk) {
            if (kfast) {
                sr = bk + 4; // 12: = 4 mod 8 (16 B elements), in {4,12} mod 16 (8 B) and in {12,20} mod 32 (4 B)
                sk = 1;
            } else {
                sk = rows + (esize == 16 ? 2 : esize == 8 ? 4 : 8); // = 2 mod 8, 4 mod 16, 8 mod 32 elements
                sr = 1;
            }
        }

        template <typename T, int BN, int BK, int STAGES, int MINB>
        void launch_mma(ContractParams p, const double *alpha, const void *v0, const void *v1,
                        const double *beta, void *vr, int device, cudaStream_t stream,
                        std::string *describe) {
            using A = typename Acc<T>::type;
            p.bn = BN;
            p.mtiles = (int)((p.M.vol + BM - 1) / BM);
            p.ntiles = (int)((p.N.vol + BN - 1) / BN);
            p.ksteps = (int)((p.K.vol + BK - 1) / BK);
            // operand enumeration: follow the contiguous direction in global memory
            p.a_kfast = p.K.n > 0 && (p.M.n == 0 || p.K.s0[0] <= p.M.s0[0]);
            p.b_kfast = p.K.n > 0 && (p.N.n == 0 || p.K.s1[0] <= p.N.s1[0]);
            tile_layout(p.a_kfast, sizeof(T), BM, BK, p.a_sr, p.a_sk);
            tile_layout(p.b_kfast, sizeof(T), BN, BK, p.b_sr, p.b_sk);
            const int a_stage = p.a_kfast ? BM * p.a_sr : BK * p.a_sk;
            const int b_stage = p.b_kfast ? BN * p.b_sr : BK * p.b_sk;
            const size_t smem = (size_t)(a_stage + b_stage) * STAGES * sizeof(T);

            // K split: fill the machine for several waves, keep every slice long
            const long long tiles = p.T.vol * p.mtiles * p.ntiles;
            const long long slots = (long long)sm_count(device, describe != nullptr) * MINB;
            int best = 1;
            double best_eff = -1;
            const int smax = std::max(1, p.ksteps / 32);
            for (int s = 1; s <= smax && s <= 1024; ++s) {
                const long long ctas = tiles * s;
                const long long waves = (ctas + slots - 1) / slots;
                double eff = (double)ctas / (double)(waves * slots);
                if (waves >= 6) eff += 1e-3; // prefer enough waves to hide the ragged tail
                if (eff > best_eff + 1e-9) best_eff = eff, best = s;
                if (waves >= 16) break;
            }
            p.ksplit = best;
            const long long ctas = tiles * p.ksplit;
            if (ctas >= (1ll << 31)) throw std::runtime_error("contraction: grid too large");
            if (describe) {
                std::stringstream ss;
                ss << "mma f64" << (sizeof(A) != sizeof(T) ? " (float operands)" : "") << " tile=" << BM << "x" << BN << "x" << BK << " stages=" << STAGES
                   << " T=" << p.T.vol << " M=" << p.M.vol << " N=" << p.N.vol << " K=" << p.K.vol
                   << " ksplit=" << p.ksplit << " ctas=" << ctas << " smem=" << smem
                   << " a_kfast=" << p.a_kfast << " b_kfast=" << p.b_kfast << " loader=cp.async";
                *describe = ss.str();
                return;
            }
            const size_t ws_bytes = (size_t)ctas * BM * BN * sizeof(A);
            A *ws = (A *)pool_alloc(device, ws_bytes);
            static size_t attr_smem[64] = {0};
            if (smem > attr_smem[device]) {
                cuda_check(cudaFuncSetAttribute(contract_mma_kernel<T, BN, BK, STAGES, MINB>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                           "cudaFuncSetAttribute");
                attr_smem[device] = smem;
            }
            {
                KernelTimer timer("contract_mma", stream);
                contract_mma_kernel<T, BN, BK, STAGES, MINB><<<(unsigned)ctas, MMA_THREADS, smem, stream>>>(
                    p, (const T *)v0, (const T *)v1, ws);
            }
            count_launch();
            cuda_check(cudaGetLastError(), "contract_mma_kernel launch");
            const long long total = p.T.vol * p.M.vol * p.N.vol;
            const unsigned grid =
                (unsigned)std::min<long long>((total + 255) / 256, (long long)sm_count(device) * 8);
            contract_reduce_kernel<T><<<grid, 256, 0, stream>>>(p, ws, (T *)vr, scalar_of<A>(alpha),
                                                               scalar_of<A>(beta));
            count_launch();
            cuda_check(cudaGetLastError(), "contract_reduce_kernel launch");
            pool_free(device, ws);
        }

    } // namespace

    int sm_count(int device, bool describe_only) {
        static int sms[64] = {0};
        if (device < 0 || device >= 64) throw std::runtime_error("invalid device");
        if (!sms[device]) {
            int n = 0;
            const cudaError_t e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
            if (e != cudaSuccess && describe_only) {
                cudaGetLastError();
                return 148; // B200; not remembered: a real launch must ask the device
            }
            cuda_check(e, "cudaDeviceGetAttribute");
            sms[device] = n;
        }
        return sms[device];
    }

    void contract(const sbk_contract_desc &desc, int dtype, const double *alpha, const void *v0,
                  const void *v1, const double *beta, void *vr, int device, cudaStream_t stream,
                  std::string *describe) {
        if (dtype != SBB_F32 && dtype != SBB_F64 && dtype != SBB_C64 && dtype != SBB_C128)
            throw std::runtime_error("contraction: unsupported type");
        if (desc.nT > GD || desc.nM > GD || desc.nN > GD || desc.nK > GD || desc.nT < 0 ||
            desc.nM < 0 || desc.nN < 0 || desc.nK < 0)
            throw std::runtime_error("contraction: too many labels in a group");
        ContractParams p;
        std::memset(&p, 0, sizeof p);
        p.T = make_group(desc.T, desc.nT, 2);
        p.M = make_group(desc.M, desc.nM, 0);
        p.N = make_group(desc.N, desc.nN, 1);
        p.K = make_group(desc.K, desc.nK, 0);
        p.conj0 = desc.conj0 != 0, p.conj1 = desc.conj1 != 0;
        if (p.T.vol == 0 || p.M.vol == 0 || p.N.vol == 0) {
            if (describe) *describe = "empty";
            return;
        }
        // An empty contracted range leaves vr = beta*vr: the generic kernel handles K.vol == 0
        const char *force = std::getenv("SBB_CONTRACT_KERNEL");
        const bool f64 = dtype == SBB_F64 || dtype == SBB_C128;
        // The tensor-core kernel wants a real tile (M, N >= 8, M*N >= 256) -- or a long contraction
        // with few outputs ("inner product" shapes, e.g. m = n = 1..12, k = 49152 of the reference's
        // tests/dist.cpp): there the generic kernel has one thread per output and no parallelism
        // over K, while the split-K of the tensor-core kernel fills the machine (the rows of the
        // 64x64 tile beyond M, N are clamped re-reads of the last row: wasted DMMA work, but the
        // operands are still read once).  The forced-"mma" parity tests cover tiny M and N.
        const long long outputs = p.T.vol * p.M.vol * p.N.vol;
        bool use_mma = p.K.vol >= 64 &&
                       ((p.M.vol * p.N.vol >= 256 && p.M.vol >= 8 && p.N.vol >= 8) ||
                        (p.K.vol >= 2048 && outputs < 148ll * 2048));
        if (force && std::strcmp(force, "simt") == 0) use_mma = false;
        if (force && std::strcmp(force, "mma") == 0 && p.K.vol > 0) use_mma = true;
        // Complex float with real tiles and a long, contiguous contracted index: the tcgen05 path
        // (TF32 x 3 split, FP32 accumulation in tensor memory; kernels_contract_tc.cu).  The FP64
        // tensor-pipe kernel below stays the path for everything else (and SBB_CONTRACT_KERNEL=mma
        // selects it for the accuracy comparison in the tests).
        if (dtype == SBB_C64 && (!force || std::strcmp(force, "tc") == 0) && p.K.n == 1 &&
            p.K.s0[0] == 1 && p.K.s1[0] == 1 && p.M.n <= 1 && p.N.n <= 1 && p.T.n <= 2 &&
            p.K.vol >= 256 && p.M.vol >= 16 && p.N.vol >= 16 && p.M.vol < (1 << 30) && p.N.vol < (1 << 30)) {
            tc::Problem tp;
            std::memset(&tp, 0, sizeof tp);
            tp.nT = p.T.n;
            for (int d = 0; d < p.T.n; ++d)
                tp.T[d] = tc::Dim{p.T.size[d], p.T.s0[d], p.T.s1[d], p.T.sr[d]};
            tp.M = tc::Dim{(int)p.M.vol, p.M.n ? p.M.s0[0] : 0, 0, p.M.n ? p.M.sr[0] : 0};
            tp.N = tc::Dim{(int)p.N.vol, 0, p.N.n ? p.N.s1[0] : 0, p.N.n ? p.N.sr[0] : 0};
            tp.K = p.K.vol, tp.conj0 = p.conj0, tp.conj1 = p.conj1;
            if (describe ? true : tc::eligible(tp, v0, v1)) {
                if (!describe || tc::eligible(tp, (const void *)16, (const void *)16)) {
                    tc::launch_c64(tp, alpha, v0, v1, beta, vr, device, stream, describe);
                    return;
                }
            }
        }
        // Small-tile regimes (validated on a B200 in round 2; bodies also checked on the CPU by
        // tests/test_row_kernel_emulation.py): long contractions whose free groups are both small go
        // to the dot kernel, short contractions with one small free group to the row kernel.
        // SBB_CONTRACT_KERNEL = simt | mma | tc | row | dot forces a path (tests).
        const bool want_dot = force ? std::strcmp(force, "dot") == 0 : true;
        const bool want_row = force ? std::strcmp(force, "row") == 0 : true;
        if (want_dot && (force || p.K.vol >= 1024) && dotk::eligible(desc)) {
            dotk::DotParams dp;
            dotk::build(desc, dp, (long long)sm_count(device, describe != nullptr) * 2048);
            // (bounds the workspace: 2 bytes per thread with the CTA tree, 256 without)
            if (dotk::threads_of(dp) <= (dotk::cta_tree(dp) ? (1ll << 26) : (1ll << 20))) {
                if (describe) {
                    std::stringstream ss;
                    ss << "dot T=" << dp.tvol << " M=" << dp.m << " N=" << dp.n << " K=" << dp.kvol
                       << " slices=" << dp.slices << " blocks=" << dp.pgm << "x" << dp.pgn;
                    *describe = ss.str();
                    return;
                }
                switch (dtype) {
                case SBB_F32: launch_dot<float>(dp, alpha, v0, v1, beta, vr, device, stream); break;
                case SBB_F64: launch_dot<double>(dp, alpha, v0, v1, beta, vr, device, stream); break;
                case SBB_C64: launch_dot<float2>(dp, alpha, v0, v1, beta, vr, device, stream); break;
                case SBB_C128: launch_dot<double2>(dp, alpha, v0, v1, beta, vr, device, stream); break;
                }
                return;
            }
        }
        if (want_row && (force || !use_mma) && rowk::eligible(desc)) {
            rowk::RowParams rp;
            bool swapped = false;
            rowk::build(desc, rp, swapped);
            if (describe) {
                std::stringstream ss;
                ss << "row rows=" << rp.rows << " small=" << rp.ns << " K=" << rp.nk
                   << " big_operand=" << (swapped ? "v1" : "v0");
                *describe = ss.str();
                return;
            }
            switch (dtype) {
            case SBB_F32: launch_row<float>(rp, swapped, alpha, v0, v1, beta, vr, device, stream); break;
            case SBB_F64: launch_row<double>(rp, swapped, alpha, v0, v1, beta, vr, device, stream); break;
            case SBB_C64: launch_row<float2>(rp, swapped, alpha, v0, v1, beta, vr, device, stream); break;
            case SBB_C128: launch_row<double2>(rp, swapped, alpha, v0, v1, beta, vr, device, stream); break;
            }
            return;
        }
        if (use_mma) {
            // 64x64 tile, BK = 8, 4 stages, 2 CTAs per SM (round 1 also measured 64x32 with 3 CTAs per SM,
            // BK = 16 with 2 stages and a cp.async.bulk row loader: all slower, removed)
#define SBB_MMA(T) launch_mma<T, 64, 8, 4, 2>(p, alpha, v0, v1, beta, vr, device, stream, describe)
            if (dtype == SBB_F64) SBB_MMA(double);
            else if (dtype == SBB_C128) SBB_MMA(double2);
            else if (dtype == SBB_F32) SBB_MMA(float);
            else SBB_MMA(float2);
#undef SBB_MMA
            return;
        }
        if (describe) {
            std::stringstream ss;
            ss << "simt T=" << p.T.vol << " M=" << p.M.vol << " N=" << p.N.vol << " K=" << p.K.vol;
            *describe = ss.str();
            return;
        }
        p.out_order[0] = 2, p.out_order[1] = 1, p.out_order[2] = 0;
        {
            // enumerate the outputs along the smallest result stride, so that the stores coalesce
            auto min_sr = [](const Group &g) {
                long long s = 1ll << 62;
                for (int d = 0; d < g.n; ++d) s = std::min(s, g.sr[d] < 0 ? -g.sr[d] : g.sr[d]);
                return s;
            };
            const long long key[3] = {min_sr(p.T), min_sr(p.M), min_sr(p.N)};
            rowk::output_order(key, p.out_order);
        }
        switch (dtype) {
        case SBB_F32: launch_simt<float>(p, alpha, v0, v1, beta, vr, device, stream); break;
        case SBB_F64: launch_simt<double>(p, alpha, v0, v1, beta, vr, device, stream); break;
        case SBB_C64: launch_simt<float2>(p, alpha, v0, v1, beta, vr, device, stream); break;
        case SBB_C128: launch_simt<double2>(p, alpha, v0, v1, beta, vr, device, stream); break;
        }
    }

} // namespace sbb
