// crc32_fold.cpp -- zlib-compatible CRC-32 (reflected polynomial 0xEDB88320) by carry-less multiplication.
//
// Host-only helper of the writer side: record checksums of checkpoint containers (ZIP local headers / central directory)
// are ON by default like the reference's torch.save produces them (async_ckpt/torch_ckpt.py:36-41 -> PyTorch's miniz
// writer, one thread, ~1 GB/s).  The table-driven slice-by-8 loop in hostbuf.cu does 1.4 GB/s per thread; folding 64 bytes
// per iteration with PCLMULQDQ (V. Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ
// Instruction", Intel 2009; constants for this polynomial as published there) does ~10 GB/s per thread, so 16 threads sum a
// 16 GB snapshot at memory speed.  Selected at run time (__builtin_cpu_supports); every other CPU keeps the table loop.
// Checked against zlib on random lengths / alignments / seeds in tests/test_hostbuf_cpu.py.
#include <stddef.h>
#include <stdint.h>

#if defined(__x86_64__)
#include <immintrin.h>

namespace {

// buf 16-byte aligned or not (unaligned loads), len >= 64 and len % 16 == 0; crc in the pre-/post-inverted domain of the caller
__attribute__((target("pclmul,sse4.1"))) uint32_t fold_pclmul(uint32_t crc, const uint8_t* buf, size_t len) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ull, 0x01c6e41596ull};  // x^(4*128+32), x^(4*128-32) mod P
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ull, 0x00ccaa009eull};  // x^(128+32),   x^(128-32)   mod P
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ull, 0x0000000000ull};  // x^64 mod P
    alignas(16) static const uint64_t poly[2] = {0x01db710641ull, 0x01f7011641ull};  // P, floor(x^64 / P)

    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x00));
    x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x10));
    x3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x20));
    x4 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128(static_cast<int>(crc)));
    x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(k1k2));
    buf += 64;
    len -= 64;

    while (len >= 64) {  // four independent 128-bit lanes, each folded over 512 bits
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00);
        x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11);
        x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x00));
        y6 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x10));
        y7 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x20));
        y8 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5);
        x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7);
        x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64;
        len -= 64;
    }

    // four lanes -> one
    x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(k3k4));
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);

    while (len >= 16) {  // remaining whole 16-byte blocks
        x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf));
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16;
        len -= 16;
    }

    // 128 -> 64 bits
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(k5k0));
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);

    // Barrett reduction to 32 bits
    x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(poly));
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return static_cast<uint32_t>(_mm_extract_epi32(x1, 1));
}

}  // namespace

extern "C" __attribute__((visibility("hidden"))) int nvrx_crc32_fold_available(void) {
    static const int ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return ok;
}

// CRC state `crc` is the raw register (already inverted by the caller); consumes the largest prefix of [p, p+n) that is a
// multiple of 16 bytes and at least 64 bytes long, returns the number of bytes consumed (0 if the range is too short).
extern "C" __attribute__((visibility("hidden"))) size_t nvrx_crc32_fold(uint32_t* crc, const uint8_t* p, size_t n) {
    const size_t take = n & ~static_cast<size_t>(15);
    if (take < 64 || !nvrx_crc32_fold_available()) return 0;
    *crc = fold_pclmul(*crc, p, take);
    return take;
}

#else

extern "C" __attribute__((visibility("hidden"))) int nvrx_crc32_fold_available(void) { return 0; }
extern "C" __attribute__((visibility("hidden"))) size_t nvrx_crc32_fold(uint32_t*, const uint8_t*, size_t) { return 0; }

#endif
