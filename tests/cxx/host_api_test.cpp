// Host-only checks of the C++ drop-in front end (include/superbblas.h): everything here runs
// without a GPU (partition generators, make_hole and the detail:: range helpers that the
// reference's tests/dist.cpp uses), compared with brute-force enumeration of the lattice sites.
// Built and run by tests/test_capi.py.
#include "superbblas.h"
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <sstream>
#include <string>

using namespace superbblas;
using namespace superbblas::detail;

#define CHECK(x)                                                                                   \
    do {                                                                                           \
        if (!(x)) {                                                                                \
            std::printf("FAILED: %s (line %d)\n", #x, __LINE__);                                   \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

template <std::size_t N> static bool inside(const Coor<N> &c, const Coor<N> &from, const Coor<N> &size, const Coor<N> &dim) {
    for (std::size_t k = 0; k < N; ++k)
        if (((c[k] - from[k]) % dim[k] + dim[k]) % dim[k] >= size[k]) return false;
    return true;
}

template <std::size_t N> static std::set<long> sites(const From_size<N> &fs, const Coor<N> &dim, bool &overlap) {
    std::set<long> s;
    overlap = false;
    for (const auto &b : fs) {
        std::size_t vol = volume<N>(b[1]);
        for (std::size_t i = 0; i < vol; ++i) {
            long idx = 0, stride = 1;
            std::size_t r = i;
            for (std::size_t k = 0; k < N; ++k) {
                int c = (int)(r % b[1][k]);
                r /= b[1][k];
                idx += ((b[0][k] + c) % dim[k]) * stride;
                stride *= dim[k];
            }
            if (!s.insert(idx).second) overlap = true;
        }
    }
    return s;
}

int main() {
    std::mt19937 rng(7);
    auto rnd = [&](int lo, int hi) { return lo + (int)(rng() % (unsigned)(hi - lo + 1)); };
    // detail::intersection against brute force, 3 dims, wrapping ranges included
    for (int it = 0; it < 2000; ++it) {
        Coor<3> dim{rnd(1, 6), rnd(1, 6), rnd(1, 6)}, f0, s0, f1, s1;
        for (int k = 0; k < 3; ++k) {
            f0[k] = rnd(0, dim[k] - 1), s0[k] = rnd(0, dim[k]);
            f1[k] = rnd(0, dim[k] - 1), s1[k] = rnd(0, dim[k]);
        }
        From_size<3> r = intersection<3>(f0, s0, f1, s1, dim);
        bool overlap;
        std::set<long> got = sites<3>(r, dim, overlap);
        CHECK(!overlap);
        std::set<long> want;
        for (int z = 0; z < dim[2]; ++z)
            for (int y = 0; y < dim[1]; ++y)
                for (int x = 0; x < dim[0]; ++x) {
                    Coor<3> c{x, y, z};
                    if (inside<3>(c, f0, s0, dim) && inside<3>(c, f1, s1, dim))
                        want.insert(x + dim[0] * (y + dim[1] * z));
                }
        CHECK(got == want);
        CHECK(volume<3>(r) == want.size());
        // make_hole: the pieces are disjoint, inside (from,size), outside the hole, and cover the rest
        auto h = make_hole<3>(f0, s0, f1, s1, dim);
        std::set<long> hs = sites<3>(h, dim, overlap);
        CHECK(!overlap);
        std::set<long> hw;
        for (int z = 0; z < dim[2]; ++z)
            for (int y = 0; y < dim[1]; ++y)
                for (int x = 0; x < dim[0]; ++x) {
                    Coor<3> c{x, y, z};
                    if (inside<3>(c, f0, s0, dim) && !inside<3>(c, f1, s1, dim))
                        hw.insert(x + dim[0] * (y + dim[1] * z));
                }
        CHECK(hs == hw);
    }
    // known answers of the reference's tests/dist.cpp:103-125
    {
        Coor<5> dim{4, 4, 4, 4, 3};
        CHECK((partitioning_distributed_procs<5>("xyztc", dim, "xyzt", 6) == Coor<5>{3, 2, 1, 1, 1}));
        CHECK((partitioning_distributed_procs<5>("xyztc", dim, "xyzt", 7) == Coor<5>{3, 2, 1, 1, 1}));
        Coor<5> dim1{4, 4, 4, 1, 3};
        CHECK((partitioning_distributed_procs<5>("xyztc", dim1, "tzyx", 32) == Coor<5>{2, 4, 4, 1, 1}));
    }
    // basic_partitioning tiles the lattice exactly once
    for (int it = 0; it < 200; ++it) {
        Coor<3> dim{rnd(1, 7), rnd(1, 7), rnd(1, 7)}, procs{rnd(1, 3), rnd(1, 3), rnd(1, 2)};
        auto p = basic_partitioning<3>("xyz", dim, procs, "zyx");
        bool overlap;
        std::set<long> s = sites<3>(p, dim, overlap);
        CHECK(!overlap);
        CHECK(s.size() == volume<3>(dim));
    }
    // error behaviour of the front end without touching a device
    {
        bool thrown = false;
        try {
            Coor<2> d{2, 2};
            partitioning_distributed_procs<2>("x", d, "x", 2); // order too short
        } catch (const std::runtime_error &) { thrown = true; }
        CHECK(thrown);
    }
    // local_copy / local_contraction (signatures of the reference's tests/local.cpp): without a CUDA
    // device they must fail loudly (there is no CPU compute path); with one, give the right answer
    {
        Coor<2> d{2, 3}, dt{3, 2};
        std::vector<double> x{1, 2, 3, 4, 5, 6}, y(6, -1.0), z(4, -1.0);
        Context cpu = createCpuContext();
        bool thrown = false;
        try {
            local_copy<2, 2, double, double>(2.0, "xy", Coor<2>{}, d, d, x.data(), nullptr, cpu, "yx",
                                             Coor<2>{}, dt, y.data(), nullptr, cpu, FastToSlow, Copy);
            Coor<2> dr{2, 2};
            local_contraction<2, 2, 2, double>(1.0, "xy", d, false, x.data(), "zy", d, false, x.data(),
                                               0.0, "xz", dr, z.data(), cpu, FastToSlow);
        } catch (const std::runtime_error &e) {
            thrown = true;
            std::printf("no device: %s\n", e.what());
        }
        if (!thrown) {
            // y[j + 3 i] = 2 x[i + 2 j];  z[i + 2 k] = sum_j x[i + 2 j] x[k + 2 j]
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 3; ++j) CHECK(y[j + 3 * i] == 2 * x[i + 2 * j]);
            for (int i = 0; i < 2; ++i)
                for (int k = 0; k < 2; ++k) {
                    double acc = 0;
                    for (int j = 0; j < 3; ++j) acc += x[i + 2 * j] * x[k + 2 * j];
                    CHECK(z[i + 2 * k] == acc);
                }
        }
    }
    // reports of the reference (performance.h:357-518): silent unless SB_TRACK_TIME / SB_TRACK_MEMORY
    // are set; with them, the calls above appear by name and nothing of the allocator is left behind
    {
        std::ostringstream os;
        reportTimings(os);
        reportCacheUsage(os);
        checkForMemoryLeaks(os);
        const char *tt = std::getenv("SB_TRACK_TIME"), *tm = std::getenv("SB_TRACK_MEMORY");
        const bool time_on = tt && std::atoi(tt) != 0, mem_on = tm && std::atoi(tm) != 0;
        const std::string text = os.str();
        CHECK((text.find("copy : ") != std::string::npos) == time_on);
        CHECK((text.find("Cache usage") != std::string::npos) == mem_on);
        resetTimings();
        std::ostringstream os2;
        reportTimings(os2);
        CHECK(os2.str().find("copy : ") == std::string::npos);
        std::printf("%s", text.c_str());
    }
    std::printf("host api ok\n");
    return 0;
}
