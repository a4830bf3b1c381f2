// tests/c/pack_team_check.cpp -- the packer team (modal_client_b200/csrc/b200pack_team.h) on its own: jobs that pull
// grains from a shared counter, as pack_parallel's do, run over and over with changing team sizes; every grain must be
// done exactly once per job and nothing may run after run() has returned.  Built twice by tests/test_c_host.py: plain,
// and with -fsanitize=thread (the data-race check for the hand-off between the caller and the parked threads).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../modal_client_b200/csrc/b200pack_team.h"

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 3000;
    b200h::PackTeam team;
    unsigned seed = 12345;
    auto rnd = [&] { seed = seed * 1103515245u + 12345u; return (seed >> 16) & 0x7fff; };
    for (int r = 0; r < rounds; ++r) {
        const int use = 1 + (int)(rnd() % 9);             // 1..9 threads (the team grows on demand)
        const unsigned grains = 1 + rnd() % 64;
        std::vector<int> done(grains, 0);                  // plain ints: a grain touched twice or concurrently is a bug
        std::atomic<unsigned> next{0};
        std::atomic<int> inside{0};
        long sum = 0;                                       // written by whoever does grain 0
        team.run(use, nullptr, [&] {
            ++inside;
            for (;;) {
                const unsigned g = next.fetch_add(1, std::memory_order_relaxed);
                if (g >= grains) break;
                done[g] += 1;
                if (g == 0) sum = 7 + (long)grains;
                if ((g & 7) == 3) std::this_thread::yield();
            }
            --inside;
        });
        if (inside.load() != 0) { fprintf(stderr, "round %d: a thread is still inside the job\n", r); return 1; }
        if (sum != 7 + (long)grains) { fprintf(stderr, "round %d: grain 0 lost\n", r); return 1; }
        for (unsigned g = 0; g < grains; ++g)
            if (done[g] != 1) { fprintf(stderr, "round %d: grain %u done %d times\n", r, g, done[g]); return 1; }
    }
    {
        b200h::PackTeam idle;  // a team that was never used, and one destroyed right after use, must shut down cleanly
        b200h::PackTeam once;
        std::atomic<int> n{0};
        once.run(4, nullptr, [&] { ++n; });
        if (n.load() < 1 || n.load() > 4) return 1;
    }
    puts("PACK TEAM OK");
    return 0;
}
