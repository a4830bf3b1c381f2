// tools/packsplit_bench.cpp -- static split vs grains from a shared counter for the slot-filling copy, with C concurrent
// "contexts" of T threads each (the map pump keeps two windows in flight; ranks share a socket).  Uses the library's
// b200h_stream_copy; no GPU.   g++ -O2 -pthread tools/packsplit_bench.cpp -o tools/packsplit_bench -ldl // static partition vs dynamic grains for the slot-filling copy, with C concurrent "contexts" of T threads each// static partition vs dynamic grains for the slot-filling copy, with C concurrent "contexts" of T threads each tools/packsplit_bench T C {0|1}
#include <dlfcn.h>
#include <sys/mman.h>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
using copy_t = void (*)(void*, const void*, size_t);
int main(int argc, char** argv) {
    void* h = dlopen(getenv("B200H_LIB") ? getenv("B200H_LIB") : "modal_client_b200/libb200hash.so", RTLD_NOW);
    copy_t copy = reinterpret_cast<copy_t>(dlsym(h, "b200h_stream_copy"));
    const int T = atoi(argv[1]), C = atoi(argv[2]), dyn = atoi(argv[3]);
    const size_t MSG = 262144, NMSG = 8192, SLOT = size_t(256) << 20, PER_SLOT = SLOT / MSG;
    auto* src = static_cast<uint8_t*>(mmap(nullptr, MSG * NMSG + (1 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    memset(src, 1, MSG * NMSG + (1 << 20));
    std::vector<uint8_t*> dsts(C);
    for (int c = 0; c < C; ++c) { dsts[c] = static_cast<uint8_t*>(mmap(nullptr, SLOT, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0)); memset(dsts[c], 0, SLOT); }
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> ctxs;
        for (int c = 0; c < C; ++c) ctxs.emplace_back([=] {
            for (size_t slot = 0; slot < NMSG / PER_SLOT; ++slot) {
                uint8_t* d = dsts[c];
                std::vector<std::thread> th;
                std::atomic<size_t> next{0};
                const size_t grain = 8;  // messages per grain (2 MiB)
                for (int k = 0; k < T; ++k) th.emplace_back([=, &next] {
                    if (!dyn) {
                        for (size_t m = PER_SLOT * k / T; m < PER_SLOT * size_t(k + 1) / T; ++m)
                            copy(d + m * MSG, src + (((slot * PER_SLOT + m) * 7919) % NMSG) * MSG + 32, MSG);
                    } else {
                        for (;;) {
                            const size_t g = next.fetch_add(1);
                            if (g * grain >= PER_SLOT) break;
                            for (size_t m = g * grain; m < (g + 1) * grain && m < PER_SLOT; ++m)
                                copy(d + m * MSG, src + (((slot * PER_SLOT + m) * 7919) % NMSG) * MSG + 32, MSG);
                        }
                    }
                });
                for (auto& x : th) x.join();
            }
        });
        for (auto& x : ctxs) x.join();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        const double g = double(MSG * NMSG) * C / dt / double(1 << 30);
        best = g > best ? g : best;
    }
    printf("%d contexts x %d threads, %s: %.2f GiB/s total\n", C, T, dyn ? "dynamic 2 MiB grains" : "static partition", best);
}
