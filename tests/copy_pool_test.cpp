// copy_pool_test.cpp — CPU exercise of interpn_b200/csrc/copy_pool.h (the threads that stage pageable caller memory for the
// host executor): sizes around the piece and vector boundaries, unaligned ends, several arrays per job, several callers
// at once. Test infrastructure: compiled and run by tests/test_copy_pool.py (g++ -O2 -pthread). Prints "OK <copies>".
#include "../interpn_b200/csrc/copy_pool.h"

#include <cstdio>

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() {
    uint64_t z = (s_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

int main() {
    using ib200::CopyPool;
    CopyPool& pool = CopyPool::get();
    const size_t cap = (size_t(9) << 20) + 4096;
    std::vector<unsigned char> src(cap * 3), dst(cap * 3 + 64);
    for (size_t i = 0; i < src.size(); i += 8) {
        const uint64_t v = rnd();
        memcpy(&src[i], &v, std::min<size_t>(8, src.size() - i));
    }
    long copies = 0;
    const size_t sizes[] = {0, 1, 15, 16, 17, 63, 64, 65, 4095, 4096, 4097, 524287, 524288, 524289, 1048576 + 3, (size_t(8) << 20), (size_t(8) << 20) + 1234567 % 4099};
    for (size_t bytes : sizes) {
        for (int soff = 0; soff < 3; ++soff) {
            for (int doff : {0, 1, 8, 24}) {
                for (int k = 1; k <= 3; ++k) {
                    std::fill(dst.begin(), dst.end(), 0xAB);
                    void* d[3];
                    const void* s[3];
                    for (int j = 0; j < k; ++j) {
                        d[j] = dst.data() + j * cap + doff;
                        s[j] = src.data() + j * cap + soff;
                    }
                    pool.copy_many(k, d, s, bytes);
                    for (int j = 0; j < k; ++j) {
                        if (memcmp(d[j], s[j], bytes) != 0) { printf("FAIL content bytes=%zu k=%d j=%d\n", bytes, k, j); return 1; }
                        const unsigned char* e = static_cast<unsigned char*>(d[j]) + bytes;
                        if (e[0] != 0xAB || (doff && static_cast<unsigned char*>(d[j])[-1] != 0xAB)) { printf("FAIL overrun bytes=%zu\n", bytes); return 1; }
                    }
                    ++copies;
                }
            }
        }
    }
    // several callers at once (one worker thread per GPU in the executor)
    std::atomic<int> bad{0};
    std::vector<std::thread> callers;
    for (int t = 0; t < 4; ++t) {
        callers.emplace_back([&, t] {
            std::vector<unsigned char> mine(size_t(5) << 20);
            for (int rep = 0; rep < 20; ++rep) {
                const size_t off = (size_t(t) * 777 + rep * 4097) % 100000, bytes = (size_t(3) << 20) + rep * 12345 + t;
                CopyPool::get().copy(mine.data() + (rep & 7), src.data() + off, bytes);
                if (memcmp(mine.data() + (rep & 7), src.data() + off, bytes) != 0) bad.fetch_add(1);
            }
        });
    }
    for (auto& c : callers) c.join();
    if (bad.load()) { printf("FAIL concurrent callers: %d\n", bad.load()); return 1; }
    printf("OK %ld copies, %d pool threads\n", copies + 80, pool.threads());
    return 0;
}
