// CPU stress of dip::CopyPool (csrc/copy_pool.h): every task of every run executes exactly once, also when
// runs follow each other faster than late workers wake up, with varying task counts and several caller threads.
#include <cstdio>
#include <vector>

#include "copy_pool.h"

int main() {
    dip::CopyPool& pool = dip::CopyPool::get();
    const unsigned width = pool.width();
    std::atomic<long> errors{0};
    auto caller = [&](unsigned seed, int rounds) {
        std::vector<std::atomic<int>> hits(64);
        unsigned x = seed;
        for (int r = 0; r < rounds; ++r) {
            x = x * 1664525u + 1013904223u;
            const unsigned n = 1 + (x >> 24) % width;
            for (auto& h : hits) h.store(0, std::memory_order_relaxed);
            long sum = 0;
            std::atomic<long> acc{0};
            for (unsigned t = 0; t < n; ++t) sum += (long)(t + 1) * (r + 1);
            pool.run(n, [&, r](unsigned t) {
                hits[t].fetch_add(1, std::memory_order_relaxed);
                acc.fetch_add((long)(t + 1) * (r + 1), std::memory_order_relaxed);
            });
            for (unsigned t = 0; t < 64; ++t)
                if (hits[t].load(std::memory_order_relaxed) != (t < n ? 1 : 0)) errors.fetch_add(1);
            if (acc.load() != sum) errors.fetch_add(1);
            if ((x & 0xFF) == 0) std::this_thread::sleep_for(std::chrono::microseconds(700));  // let the workers fall asleep
        }
    };
    std::thread a(caller, 1u, 60000), b(caller, 2u, 60000), c(caller, 3u, 60000);
    caller(4u, 60000);
    a.join(); b.join(); c.join();
    std::printf("width %u, errors %ld\n", width, errors.load());
    return errors.load() ? 1 : 0;
}
