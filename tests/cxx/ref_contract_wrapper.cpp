// Runs the reference's OWN contraction test-suite (tests/contract.cpp of eromero-vlc/superbblas,
// included from where it lies under /root/reference — never copied) against THIS library through the
// drop-in header include/superbblas.h.  The reference's `main` runs ~1.2 million contractions, twice;
// this wrapper keeps the test templates untouched and selects which part of the enumeration to run:
//
//   ref_contract_wrapper [--nt=0|1] [--type=s|d|c|z] [--components=N] [--cpu]
//
// --nt picks the number of batch labels (the outermost enumeration level, test_for_A<NT,T>),
// --cpu passes host (CPU-context) components instead of GPU ones (they are staged through the GPU).
#define main reference_main
#include "tests/contract.cpp"
#undef main

int main(int argc, char **argv) {
    int nt = 1, ncomponents = 1;
    char type = 'z';
    bool on_cpu = false;
    for (int i = 1; i < argc; ++i) {
        if (std::strncmp("--nt=", argv[i], 5) == 0) nt = std::atoi(argv[i] + 5);
        else if (std::strncmp("--type=", argv[i], 7) == 0) type = argv[i][7];
        else if (std::strncmp("--components=", argv[i], 13) == 0) ncomponents = std::atoi(argv[i] + 13);
        else if (std::strcmp("--cpu", argv[i]) == 0) on_cpu = true;
    }
    initialize_test();
    if (on_cpu) {
        std::vector<Context> ctx(ncomponents, createCpuContext());
        std::vector<superbblas::detail::Cpu> xpus;
        for (const auto &i : ctx) xpus.push_back(i.toCpu(0));
        if (type == 'd') nt ? test_for_A<1, double>(ctx, xpus) : test_for_A<0, double>(ctx, xpus);
        else if (type == 's') nt ? test_for_A<1, float>(ctx, xpus) : test_for_A<0, float>(ctx, xpus);
        else if (type == 'c') nt ? test_for_A<1, std::complex<float>>(ctx, xpus) : test_for_A<0, std::complex<float>>(ctx, xpus);
        else nt ? test_for_A<1, std::complex<double>>(ctx, xpus) : test_for_A<0, std::complex<double>>(ctx, xpus);
    } else {
        std::vector<Context> ctx;
        for (int i = 0; i < ncomponents; ++i) ctx.push_back(createGpuContext(i % getGpuDevicesCount()));
        std::vector<superbblas::detail::Gpu> xpus;
        for (const auto &i : ctx) xpus.push_back(i.toGpu(0));
        if (type == 'd') nt ? test_for_A<1, double>(ctx, xpus) : test_for_A<0, double>(ctx, xpus);
        else if (type == 's') nt ? test_for_A<1, float>(ctx, xpus) : test_for_A<0, float>(ctx, xpus);
        else if (type == 'c') nt ? test_for_A<1, std::complex<float>>(ctx, xpus) : test_for_A<0, std::complex<float>>(ctx, xpus);
        else nt ? test_for_A<1, std::complex<double>>(ctx, xpus) : test_for_A<0, std::complex<double>>(ctx, xpus);
    }
    clearCaches();
    clearHandles();
    std::cout << " Everything went ok! (" << test_number << " contractions of the reference suite)" << std::endl;
    return 0;
}
