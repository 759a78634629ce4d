// Stress test of csrc/gf_copy_pool.h (tests/test_copy_pool.py builds and runs it): random block sizes around the
// single-thread threshold, thread counts 1..16 changing from job to job, callers on several threads at once.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "gf_copy_pool.h"

static int run_caller(unsigned seed, int jobs)
{
    std::mt19937 rng(seed);
    const size_t cap = (size_t)24 << 20;
    std::vector<unsigned char> src(cap), dst(cap);
    for (size_t i = 0; i < cap; ++i) src[i] = (unsigned char)(rng() >> 7);
    for (int j = 0; j < jobs; ++j) {
        const int kind = (int)(rng() % 4);
        size_t n = kind == 0 ? rng() % 4096 : (kind == 1 ? ((size_t)1 << 20) - 64 + rng() % 128 : ((size_t)1 << 20) + rng() % (cap - ((size_t)1 << 20) - 4096));
        const size_t so = rng() % 2048, dof = rng() % 2048;          // unaligned starts
        if (so + n > cap) n = cap - so;
        if (dof + n > cap) n = cap - dof;
        const int threads = 1 + (int)(rng() % 16);
        std::memset(dst.data(), 0xA5, dst.size() < n + dof + 64 ? dst.size() : n + dof + 64);
        GfCopyPool::get().copy(dst.data() + dof, src.data() + so, n, threads);
        if (std::memcmp(dst.data() + dof, src.data() + so, n) != 0) { std::printf("MISMATCH seed %u job %d n %zu threads %d\n", seed, j, n, threads); return 1; }
        if (dof > 0 && dst[dof - 1] != 0xA5) { std::printf("UNDERRUN seed %u job %d\n", seed, j); return 1; }
        if (dof + n < cap && dst[dof + n] != 0xA5) { std::printf("OVERRUN seed %u job %d n %zu threads %d\n", seed, j, n, threads); return 1; }
    }
    return 0;
}

int main(int argc, char** argv)
{
    const int jobs = argc > 1 ? std::atoi(argv[1]) : 300;
    int rc[3] = {0, 0, 0};
    std::vector<std::thread> callers;
    for (int c = 0; c < 3; ++c) callers.emplace_back([&, c] { rc[c] = run_caller(1234u + 77u * c, jobs); });
    for (auto& t : callers) t.join();
    if (rc[0] || rc[1] || rc[2]) return 1;
    std::printf("copy pool ok: 3 callers x %d jobs\n", jobs);
    return 0;
}
