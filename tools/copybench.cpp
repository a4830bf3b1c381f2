// tools/copybench.cpp -- the staging copy (b200h_stream_copy, csrc/b200pack_copy.cpp) the way pack_parallel uses it:
// T threads fill 256 MiB slots from 16 384 scattered 256 KiB sources (4 GiB per pass), best of 4 passes.  Needs no GPU
// (the library is only dlopen'ed for the symbol).  B200H_COPY_ISA=avx2|plain caps the instruction set.
//   g++ -O2 -pthread tools/copybench.cpp -o tools/copybench -ldl && tools/copybench 8
#include <dlfcn.h>
#include <sys/mman.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

int main(int argc, char** argv) {
    const char* lib = getenv("B200H_LIB") ? getenv("B200H_LIB") : "modal_client_b200/libb200hash.so";
    void* h = dlopen(lib, RTLD_NOW);
    if (!h) {
        fprintf(stderr, "%s\n", dlerror());
        return 1;
    }
    auto copy = reinterpret_cast<void (*)(void*, const void*, size_t)>(dlsym(h, "b200h_stream_copy"));
    auto isa = reinterpret_cast<const char* (*)()>(dlsym(h, "b200h_stream_copy_isa"));
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t MSG = 262144, NMSG = 16384, SLOT = size_t(256) << 20, PER_SLOT = SLOT / MSG;
    auto* src = static_cast<uint8_t*>(mmap(nullptr, MSG * NMSG + (1 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    auto* dst = static_cast<uint8_t*>(mmap(nullptr, 2 * SLOT, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    memset(src, 1, MSG * NMSG + (1 << 20));
    memset(dst, 0, 2 * SLOT);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t c = 0; c < NMSG / PER_SLOT; ++c) {  // one pinned slot after the other, threads spawned per slot
            uint8_t* d = dst + (c & 1) * SLOT;
            std::vector<std::thread> th;
            for (int k = 0; k < T; ++k)
                th.emplace_back([=] {
                    for (size_t m = PER_SLOT * k / T; m < PER_SLOT * size_t(k + 1) / T; ++m)  // +32: a bytes object's header
                        copy(d + m * MSG, src + (((c * PER_SLOT + m) * 7919) % NMSG) * MSG + 32, MSG);
                });
            for (auto& x : th) x.join();
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double g = double(MSG * NMSG) / dt / double(1 << 30);
        best = g > best ? g : best;
    }
    printf("%s, %d threads: %.2f GiB/s\n", isa(), T, best);
    return 0;
}
