// Host staging copy rates: glibc memcpy vs explicit non-temporal (streaming) stores, T threads, one 256 MB chunk.
// build: g++ -O2 -pthread -o host_copy host_copy.cpp ; run: ./host_copy [threads]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <emmintrin.h>
#include <thread>
#include <vector>

static void nt_copy(char* d, const char* s, size_t n)
{
    while (n && (reinterpret_cast<uintptr_t>(d) & 15)) { *d++ = *s++; --n; }
    size_t k = n / 64;
    for (size_t i = 0; i < k; ++i) {
        __m128i a = _mm_loadu_si128((const __m128i*)(s)), b = _mm_loadu_si128((const __m128i*)(s + 16));
        __m128i c = _mm_loadu_si128((const __m128i*)(s + 32)), e = _mm_loadu_si128((const __m128i*)(s + 48));
        _mm_stream_si128((__m128i*)d, a); _mm_stream_si128((__m128i*)(d + 16), b);
        _mm_stream_si128((__m128i*)(d + 32), c); _mm_stream_si128((__m128i*)(d + 48), e);
        s += 64; d += 64;
    }
    _mm_sfence();
    memcpy(d, s, n - k * 64);
}
template <typename F> static double run(F&& f, char* d, const char* s, size_t n, int T)
{
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { size_t a = n * t / T, b = n * (t + 1) / T; f(d + a, s + a, b - a); });
        for (auto& x : th) x.join();
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt < best) best = dt;
    }
    return n / best / 1e9;
}
int main(int argc, char** argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t n = (size_t)256 << 20;
    char* s = (char*)aligned_alloc(4096, n); char* d = (char*)aligned_alloc(4096, n);
    memset(s, 1, n); memset(d, 2, n);
    printf("{\"threads\": %d, \"memcpy_GBs\": %.1f, \"nt_store_GBs\": %.1f}\n", T,
           run([](char* a, const char* b, size_t m) { memcpy(a, b, m); }, d, s, n, T), run(nt_copy, d, s, n, T));
    return 0;
}
