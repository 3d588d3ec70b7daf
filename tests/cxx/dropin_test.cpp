// Drop-in check of include/superbblas.h: a caller written against the reference's public API
// (same calls as SURVEY Appendix A / the reference's tests/contract.cpp and tests/dist.cpp) compiled
// against this library and run on a B200.  Every result is verified on the host by brute force.
#include "superbblas.h"
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

using namespace superbblas;
using Z = std::complex<double>;
using C = std::complex<float>;

#define CHECK(cond)                                                                                \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);                          \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

template <typename T> T *to_device(const std::vector<T> &h) {
    T *d = nullptr;
    if (cudaMalloc((void **)&d, sizeof(T) * std::max<std::size_t>(h.size(), 1)) != cudaSuccess)
        std::abort();
    cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice);
    return d;
}
template <typename T> std::vector<T> to_host(const T *d, std::size_t n) {
    std::vector<T> h(n);
    cudaMemcpy(h.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost);
    return h;
}

int main() {
    if (getGpuDevicesCount() == 0) {
        std::printf("no GPU\n");
        return 2;
    }
    Context gpu = createGpuContext(0), cpu = createCpuContext();
    const int L = 8, Lt = 16;

    // --- config 1: permuting copy "xyztsc" -> "cstzyx", one component ------------------------------
    {
        Coor<6> dim0{L, L, L, Lt, 4, 3}, dim1{3, 4, Lt, L, L, L};
        PartitionItem<6> p0{Coor<6>{}, dim0}, p1{Coor<6>{}, dim1};
        std::size_t vol = detail::volume(dim0);
        std::vector<Z> h0(vol);
        for (std::size_t i = 0; i < vol; ++i) h0[i] = Z((double)i, -(double)i);
        Z *d0 = to_device(h0), *d1 = to_device(std::vector<Z>(vol));
        const Z *src = d0;
        copy<6, 6>(Z{1}, &p0, 1, "xyztsc", {}, dim0, dim0, &src, nullptr, &gpu, &p1, 1, "cstzyx", {},
                   dim1, &d1, nullptr, &gpu, FastToSlow, Copy);
        sync(gpu);
        std::vector<Z> h1 = to_host(d1, vol);
        std::size_t bad = 0;
        for (int x = 0; x < L; ++x)
            for (int y = 0; y < L; ++y)
                for (int z = 0; z < L; ++z)
                    for (int t = 0; t < Lt; ++t)
                        for (int s = 0; s < 4; ++s)
                            for (int c = 0; c < 3; ++c) {
                                std::size_t i0 = x + L * (y + L * (z + L * (t + Lt * (s + 4 * c))));
                                std::size_t i1 = c + 3 * (s + 4 * (t + Lt * (z + L * (y + L * x))));
                                if (h1[i1] != h0[i0]) ++bad;
                            }
        CHECK(bad == 0);

        // contraction over s,c: R[x,y,z,t] = sum_{s,c} conj(A) B
        std::vector<Z> hb(vol);
        for (std::size_t i = 0; i < vol; ++i) hb[i] = Z(1.0 / (1 + i % 7), 0.25 * (i % 5));
        Z *db = to_device(hb);
        Coor<4> dimr{L, L, L, Lt};
        PartitionItem<4> pr{Coor<4>{}, dimr};
        Z *dr = to_device(std::vector<Z>(detail::volume(dimr), Z{1, 1}));
        const Z *a = d0, *b = db;
        contraction<6, 6, 4>(Z{1}, &p0, {}, dim0, dim0, 1, "xyztsc", true, &a, &gpu, &p0, {}, dim0,
                             dim0, 1, "xyztsc", false, &b, &gpu, Z{0}, &pr, {}, dimr, dimr, 1, "xyzt",
                             &dr, &gpu, FastToSlow);
        sync(gpu);
        std::vector<Z> hr = to_host(dr, detail::volume(dimr));
        std::size_t site = (std::size_t)L * L * L * Lt;
        double maxrel = 0;
        for (std::size_t i = 0; i < site; ++i) {
            Z acc = 0;
            for (int k = 0; k < 12; ++k) acc += std::conj(h0[i + site * k]) * hb[i + site * k];
            maxrel = std::max(maxrel, std::abs(acc - hr[i]) / (std::abs(acc) + 1e-300));
        }
        CHECK(maxrel < 1e-12);
        cudaFree(d0), cudaFree(d1), cudaFree(db), cudaFree(dr);
    }

    // --- config 2 (reduced): distillation contraction R[t,n,m] = sum_k conj(V0[k,t,n]) V1[k,t,m] -------
    {
        const int Ls = 4, Lts = 4, nv = 16;
        Coor<6> dimv{3, Ls, Ls, Ls, Lts, nv};
        Coor<3> dimr{Lts, nv, nv};
        PartitionItem<6> pv{Coor<6>{}, dimv};
        PartitionItem<3> pr{Coor<3>{}, dimr};
        std::size_t K = 3 * Ls * Ls * Ls, vol = detail::volume(dimv);
        std::vector<Z> h0(vol), h1(vol);
        for (std::size_t i = 0; i < vol; ++i)
            h0[i] = Z(std::sin(0.1 * i), std::cos(0.3 * i)), h1[i] = Z(std::cos(0.7 * i), std::sin(0.2 * i));
        // host (CPU context) operands: the library stages them through the GPU
        std::vector<Z> hr(detail::volume(dimr));
        const Z *a = h0.data(), *b = h1.data();
        Z *c = hr.data();
        contraction<6, 6, 3>(Z{1}, &pv, {}, dimv, dimv, 1, "cxyztn", true, &a, &cpu, &pv, {}, dimv,
                             dimv, 1, "cxyztm", false, &b, &cpu, Z{0}, &pr, {}, dimr, dimr, 1, "tnm",
                             &c, &cpu, FastToSlow);
        double maxrel = 0;
        for (int t = 0; t < Lts; ++t)
            for (int n = 0; n < nv; ++n)
                for (int m = 0; m < nv; ++m) {
                    Z acc = 0;
                    for (std::size_t k = 0; k < K; ++k)
                        acc += std::conj(h0[k + K * (t + Lts * n)]) * h1[k + K * (t + Lts * m)];
                    maxrel = std::max(maxrel, std::abs(acc - hr[t + Lts * (n + nv * m)]) / std::abs(acc));
                }
        CHECK(maxrel < 1e-12);
    }

    // --- config 3/5 (reduced): 8 components on one GPU, t-slabs -> (z,t) blocks, then a periodic shift ----
    {
        const int P = 8;
        Coor<7> dim{4, 4, 4, 8, 4, 3, 2};
        auto p0 = basic_partitioning<7>("xyztscn", dim, Coor<7>{1, 1, 1, P, 1, 1, 1}, "t", P, 1);
        auto p1 = basic_partitioning<7>("xyztscn", dim, Coor<7>{1, 1, 2, P / 2, 1, 1, 1}, "zt", P, 1);
        std::size_t vol = detail::volume(dim);
        std::vector<C> glob(vol);
        for (std::size_t i = 0; i < vol; ++i) glob[i] = C((float)i, (float)(i % 13));
        // scatter the global tensor (one CPU component) onto the 8 GPU components of p0
        PartitionItem<7> pg{Coor<7>{}, dim};
        std::vector<C *> v0(P), v1(P);
        std::vector<Context> ctx(P, gpu);
        for (int i = 0; i < P; ++i) {
            v0[i] = to_device(std::vector<C>(detail::volume(p0[i][1])));
            v1[i] = to_device(std::vector<C>(detail::volume(p1[i][1])));
        }
        const C *g = glob.data();
        copy<7, 7>(C{1}, &pg, 1, "xyztscn", {}, dim, dim, &g, nullptr, &cpu, p0.data(), P, "xyztscn",
                   {}, dim, v0.data(), nullptr, ctx.data(), FastToSlow, Copy);
        copy<7, 7>(C{1}, p0.data(), P, "xyztscn", {}, dim, dim, (const C **)v0.data(), nullptr,
                   ctx.data(), p1.data(), P, "xyztscn", {}, dim, v1.data(), nullptr, ctx.data(),
                   FastToSlow, Copy);
        // shift by (0,0,1,1,...) while gathering back to one CPU component
        std::vector<C> back(vol);
        C *bp = back.data();
        Coor<7> from1{0, 0, 1, 1, 0, 0, 0};
        copy<7, 7>(C{1}, p1.data(), P, "xyztscn", {}, dim, dim, (const C **)v1.data(), nullptr,
                   ctx.data(), &pg, 1, "xyztscn", from1, dim, &bp, nullptr, &cpu, FastToSlow, Copy);
        std::size_t bad = 0;
        Coor<7> c;
        for (std::size_t i = 0; i < vol; ++i) {
            std::size_t r = i, j = 0, stride = 1;
            for (int k = 0; k < 7; ++k) {
                c[k] = r % dim[k];
                r /= dim[k];
                j += ((c[k] + from1[k]) % dim[k]) * stride;
                stride *= dim[k];
            }
            if (back[j] != glob[i]) ++bad;
        }
        CHECK(bad == 0);
        for (int i = 0; i < P; ++i) cudaFree(v0[i]), cudaFree(v1[i]);
    }

    // --- masked copy (MaskType masks on both tensors): even sites of a permuted field -----------------------
    {
        Coor<4> dim0{L, L, L, Lt}, dim1{Lt, L, L, L};
        PartitionItem<4> p0{Coor<4>{}, dim0}, p1{Coor<4>{}, dim1};
        std::size_t vol = detail::volume(dim0);
        std::vector<Z> h0(vol), h1(vol, Z{-5, 5});
        std::vector<MaskType> m0(vol), m1(vol);
        for (int x = 0; x < L; ++x)
            for (int y = 0; y < L; ++y)
                for (int z = 0; z < L; ++z)
                    for (int t = 0; t < Lt; ++t) {
                        std::size_t i0 = x + L * (y + L * (z + L * t));
                        std::size_t i1 = t + Lt * (z + L * (y + L * x));
                        h0[i0] = Z((double)i0, 0.5);
                        m0[i0] = m1[i1] = (x + y + z + t) % 2 == 0 ? 1.0f : 0.0f;
                    }
        Z *d0 = to_device(h0), *d1 = to_device(h1);
        MaskType *dm0 = to_device(m0), *dm1 = to_device(m1);
        const Z *src = d0;
        const MaskType *pm0 = dm0, *pm1 = dm1;
        copy<4, 4>(Z{2}, &p0, 1, "xyzt", {}, dim0, dim0, &src, &pm0, &gpu, &p1, 1, "tzyx", {}, dim1, &d1,
                   &pm1, &gpu, FastToSlow, Copy);
        sync(gpu);
        std::vector<Z> r = to_host(d1, vol);
        std::size_t bad = 0;
        for (int x = 0; x < L; ++x)
            for (int y = 0; y < L; ++y)
                for (int z = 0; z < L; ++z)
                    for (int t = 0; t < Lt; ++t) {
                        std::size_t i0 = x + L * (y + L * (z + L * t));
                        std::size_t i1 = t + Lt * (z + L * (y + L * x));
                        Z want = (x + y + z + t) % 2 == 0 ? Z{2} * h0[i0] : Z{-5, 5};
                        if (r[i1] != want) ++bad;
                    }
        CHECK(bad == 0);
        cudaFree(d0), cudaFree(d1), cudaFree(dm0), cudaFree(dm1);
    }

    // --- Request / wait (dist.h:54-61, :3554-3557): a permuting copy into a host destination is begun,
    //     completed by wait(), and the request can be copied and waited on again ---------------------------
    {
        Coor<3> d0{5, 4, 3}, d1{3, 4, 5};
        PartitionItem<3> p0{Coor<3>{}, d0}, p1{Coor<3>{}, d1};
        std::vector<double> h0(60), h1(60, -1.0);
        for (int i = 0; i < 60; ++i) h0[i] = i;
        double *dsrc = to_device(h0);
        const double *src = dsrc;
        double *dst = h1.data();
        Request req;
        copy<3, 3>(2.0, &p0, 1, "abc", {}, d0, d0, &src, nullptr, &gpu, &p1, 1, "cba", {}, d1, &dst, nullptr,
                   &cpu, FastToSlow, Copy, &req);
        Request again = req;
        wait(req);
        wait(again);
        int bad = 0;
        for (int a = 0; a < 5; ++a)
            for (int b = 0; b < 4; ++b)
                for (int c = 0; c < 3; ++c)
                    if (h1[c + 3 * (b + 4 * a)] != 2.0 * h0[a + 5 * (b + 4 * c)]) ++bad;
        CHECK(bad == 0);
        cudaFree(dsrc);
    }

    // --- error behaviour ------------------------------------------------------------------------------------
    {
        Coor<2> d{2, 2};
        PartitionItem<2> p{Coor<2>{}, d};
        std::vector<double> x(4), y(4);
        const double *xp = x.data();
        double *yp = y.data();
        bool thrown = false;
        try {
            copy<2, 2>(1.0, &p, 1, "xy", {}, d, d, &xp, nullptr, &cpu, &p, 1, "xz", {}, d, &yp,
                       nullptr, &cpu, FastToSlow, Copy);
        } catch (const std::runtime_error &e) {
            thrown = std::string(e.what()) == "Invalid copy operation";
        }
        CHECK(thrown);
    }
    clearCaches();
    std::printf("Everything went ok!\n");
    return 0;
}
