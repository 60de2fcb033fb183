// Host-side repacking rate: uint64 words (36 significant bits) <-> the 5-byte packed format (a plane of 32-bit low words + a
// plane of high bytes per limb), T threads, streaming stores.  Decides whether hml_hmult_host could pack on the CPU around its
// PCIe copies.   g++ -O3 -mavx2 -pthread host_pack_bench.cpp -o host_pack_bench && ./host_pack_bench [threads]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>

static const size_t N = 65536;
static void pack_limb(const uint64_t *w, unsigned char *limb) {
  uint32_t *lo = (uint32_t *)limb;
  unsigned char *hi = limb + 4 * N;
  for (size_t i = 0; i < N; i += 8) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(w + i)), b = _mm256_loadu_si256((const __m256i *)(w + i + 4));
    // low words: even 32-bit lanes of a and b
    __m256i pa = _mm256_permutevar8x32_epi32(a, _mm256_setr_epi32(0, 2, 4, 6, 1, 3, 5, 7));
    __m256i pb = _mm256_permutevar8x32_epi32(b, _mm256_setr_epi32(0, 2, 4, 6, 1, 3, 5, 7));
    __m256i lows = _mm256_permute2x128_si256(pa, pb, 0x20), highs = _mm256_permute2x128_si256(pa, pb, 0x31);
    _mm256_stream_si256((__m256i *)(lo + i), lows);
    // high bytes: 8 x u32 -> 8 x u8
    __m128i h16 = _mm_packus_epi32(_mm256_castsi256_si128(highs), _mm256_extracti128_si256(highs, 1));
    __m128i h8 = _mm_packus_epi16(h16, h16);
    _mm_storel_epi64((__m128i *)(hi + i), h8);
  }
}
static void unpack_limb(const unsigned char *limb, uint64_t *w) {
  const uint32_t *lo = (const uint32_t *)limb;
  const unsigned char *hi = limb + 4 * N;
  for (size_t i = 0; i < N; i += 4) {
    __m128i l = _mm_loadu_si128((const __m128i *)(lo + i));
    __m128i h = _mm_cvtepu8_epi32(_mm_cvtsi32_si128(*(const int *)(hi + i)));
    __m256i v = _mm256_or_si256(_mm256_cvtepu32_epi64(l), _mm256_slli_epi64(_mm256_cvtepu32_epi64(h), 32));
    _mm256_stream_si256((__m256i *)(w + i), v);
  }
}
int main(int argc, char **argv) {
  const int T = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
  const size_t limbs = 2 * 2 * 35 * 8;  // eight hmult inputs
  uint64_t *words = (uint64_t *)aligned_alloc(64, limbs * N * 8);
  unsigned char *packed = (unsigned char *)aligned_alloc(64, limbs * N * 5);
  for (size_t i = 0; i < limbs * N; ++i) words[i] = (i * 0x9E3779B97F4A7C15ull) >> 28;
  auto run = [&](bool pack) {
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < T; ++t)
      th.emplace_back([&, t] {
        for (size_t l = t; l < limbs; l += T) pack ? pack_limb(words + l * N, packed + l * 5 * N) : unpack_limb(packed + l * 5 * N, words + l * N);
      });
    for (auto &x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };
  run(true);
  double tp = 1e9, tu = 1e9;
  for (int r = 0; r < 5; ++r) { tp = std::min(tp, run(true)); tu = std::min(tu, run(false)); }
  // check
  uint64_t *chk = (uint64_t *)aligned_alloc(64, N * 8);
  unpack_limb(packed + 3 * 5 * N, chk);
  int ok = 1;
  for (size_t i = 0; i < N; ++i) ok &= chk[i] == ((((3 * N + i)) * 0x9E3779B97F4A7C15ull) >> 28);
  printf("threads %d: pack %.1f GB/s of words in (%.3f ms per hmult's 73.4 MB), unpack %.1f GB/s of words out (%.3f ms per 36.7 MB), roundtrip %s\n", T,
         limbs * N * 8 / tp / 1e9, 73.4e6 / (limbs * N * 8 / tp) * 1e3, limbs * N * 8 / tu / 1e9, 36.7e6 / (limbs * N * 8 / tu) * 1e3, ok ? "ok" : "MISMATCH");
  return !ok;
}
