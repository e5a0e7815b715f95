/* Experiment: the block gather of LoopDetectorBranchBound::Detect alone, on the host's cores (no GPU, staging in
 * ordinary memory). Build and run from the repo root:
 *   python scripts/exp_gather_cpu.py [threads] [groups] [shuffle]
 * (the script cuts the BlockGatherer class out of host/src/loop_detector.cpp into a scratch file next to this one) */

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>
#include <algorithm>
#include <random>
#include <emmintrin.h>
static void* csm_alloc_pinned(size_t n) { return aligned_alloc(4096, (n + 4095) & ~size_t(4095)); }
static void csm_free_pinned(void* p) { free(p); }
#include "gatherer_class.inc"
int main(int argc, char** argv)
{
    const int threads = argc > 1 ? atoi(argv[1]) : 8;
    const int groups = argc > 2 ? atoi(argv[2]) : 4;
    const int shuffle = argc > 3 ? atoi(argv[3]) : 0;
    const size_t per_group = 21760;            // 64 maps x 340 blocks
    std::vector<std::uint16_t*> blocks;
    for (size_t i = 0; i < per_group * groups; ++i) { auto* p = new std::uint16_t[256]; memset(p, 1, 512); blocks.push_back(p); }
    if (shuffle) { std::mt19937 r(1); std::shuffle(blocks.begin(), blocks.end(), r); }
    BlockGatherer g(threads);
    for (int k = 0; k < groups; ++k) g.Area(k, per_group * 512);
    for (int rep = 0; rep < 8; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        for (int k = 0; k < groups; ++k) {
            std::vector<const std::uint16_t*> src(blocks.begin() + k * per_group, blocks.begin() + (k + 1) * per_group);
            g.Start(std::move(src), (std::uint16_t*)g.Area(k, per_group * 512), 512);
            g.Finish();
        }
        auto t1 = std::chrono::steady_clock::now();
        printf("rep %d: %.3f ms\n", rep, std::chrono::duration<double, std::milli>(t1 - t0).count());
    }
}
