// Host memory bandwidth probe for the e2e bound of b200seg_postproc_soma_host_batch: T threads, each on its own 256 MB,
// (a) memcpy, (b) non-temporal fill, (c) read-only sum.  Build: g++ -O2 -pthread -o host_bw host_bw.cpp; run: ./host_bw [threads]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <emmintrin.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    const size_t N = 256u << 20;
    std::vector<char*> a(T), b(T);
    for (int t = 0; t < T; ++t) { a[t] = (char*)aligned_alloc(4096, N); b[t] = (char*)aligned_alloc(4096, N); memset(a[t], 1, N); memset(b[t], 2, N); }
    const char* names[3] = {"memcpy (read + write)", "non-temporal fill (write)", "sum (read)"};
    for (int mode = 0; mode < 3; ++mode) {
        double best = 1e9;
        for (int rep = 0; rep < 3; ++rep) {
            std::vector<std::thread> th;
            const double t0 = now();
            for (int t = 0; t < T; ++t)
                th.emplace_back([&, t] {
                    if (mode == 0) memcpy(b[t], a[t], N);
                    else if (mode == 1) { const __m128i z = _mm_set1_epi8((char)rep); for (size_t i = 0; i < N; i += 16) _mm_stream_si128((__m128i*)(b[t] + i), z); _mm_sfence(); }
                    else { uint64_t s = 0; const uint64_t* p = (const uint64_t*)a[t]; for (size_t i = 0; i < N / 8; ++i) s += p[i]; if (s == 42) printf("!"); }
                });
            for (auto& x : th) x.join();
            const double dt = now() - t0;
            if (dt < best) best = dt;
        }
        const double bytes = (double)N * T * (mode == 0 ? 2 : 1);
        printf("%d threads, %s: %.1f GB/s\n", T, names[mode], bytes / best / 1e9);
    }
    return 0;
}
