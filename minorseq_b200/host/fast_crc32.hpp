// fast_crc32.hpp -- CRC-32 (the gzip polynomial) by carry-less multiplication: after the project's own DEFLATE decoder
// the per-block CRC check of a BGZF member is a third of the decode time with zlib's table-driven crc32().  Folding four
// 128-bit lanes with PCLMULQDQ (Gopal et al., "Fast CRC Computation for Generic Polynomials Using PCLMULQDQ
// Instruction", Intel 2009) runs at memory speed.  Selected at run time; anything else goes through zlib, and
// host_selftest compares the two on random buffers.
//
// Third-party provenance: crc32_clmul below -- its folding constants (k1k2, k3k4, k5k0, poly) and the structure of the
// fold / reduce steps -- follows the well-known SSE4.2 + PCLMULQDQ CRC-32 routine distributed with Chromium's zlib
// (crc32_simd.c, crc32_sse42_simd_), Copyright 2017 The Chromium Authors, BSD-3-Clause licence
// (https://chromium.googlesource.com/chromium/src/third_party/zlib, LICENSE file of the Chromium project: redistribution
// in source and binary form permitted with this notice retained; provided "as is" without warranty; the names of the
// copyright holders may not be used to endorse derived products).  It is not derived from /root/reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include <zlib.h>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define MS_HAVE_CLMUL_CRC 1
#endif

namespace mscrc {

#ifdef MS_HAVE_CLMUL_CRC
// len: a multiple of 16, at least 64; crc: the running value in its inverted (internal) form
__attribute__((target("pclmul,sse4.1"))) inline uint32_t crc32_clmul(const unsigned char* buf, size_t len, uint32_t crc) {
    alignas(16) static const uint64_t k1k2[2] = {0x0154442bd4ULL, 0x01c6e41596ULL};
    alignas(16) static const uint64_t k3k4[2] = {0x01751997d0ULL, 0x00ccaa009eULL};
    alignas(16) static const uint64_t k5k0[2] = {0x0163cd6124ULL, 0x0000000000ULL};
    alignas(16) static const uint64_t poly[2] = {0x01db710641ULL, 0x01f7011641ULL};
    __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
    x1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x00));
    x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x10));
    x3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x20));
    x4 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x30));
    x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128(static_cast<int>(crc)));
    x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(k1k2));
    buf += 64; len -= 64;
    while (len >= 64) {             // four lanes in parallel
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
        x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
        x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
        x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
        y5 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x00));
        y6 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x10));
        y7 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x20));
        y8 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf + 0x30));
        x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
        x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
        buf += 64; len -= 64;
    }
    x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(k3k4));      // four lanes -> one
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
    while (len >= 16) {
        x2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(buf));
        x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
        buf += 16; len -= 16;
    }
    x2 = _mm_clmulepi64_si128(x1, x0, 0x10);                            // 128 -> 64 bits
    x3 = _mm_setr_epi32(~0, 0, ~0, 0);
    x1 = _mm_srli_si128(x1, 8);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_loadl_epi64(reinterpret_cast<const __m128i*>(k5k0));
    x2 = _mm_srli_si128(x1, 4);
    x1 = _mm_and_si128(x1, x3);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(poly));        // Barrett reduction to 32 bits
    x2 = _mm_and_si128(x1, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
    x2 = _mm_and_si128(x2, x3);
    x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x1 = _mm_xor_si128(x1, x2);
    return static_cast<uint32_t>(_mm_extract_epi32(x1, 1));
}
inline bool have_clmul() {
    static const bool ok = __builtin_cpu_supports("pclmul") && __builtin_cpu_supports("sse4.1");
    return ok;
}
#endif

// same value as zlib's crc32(crc32(0, NULL, 0), buf, len)
inline uint32_t crc32(const unsigned char* buf, size_t len) {
    uint32_t crc = 0;
#ifdef MS_HAVE_CLMUL_CRC
    if (len >= 64 && have_clmul()) {
        const size_t chunk = len & ~static_cast<size_t>(15);
        crc = ~crc32_clmul(buf, chunk, ~crc);
        buf += chunk; len -= chunk;
        if (!len) return crc;
    }
#endif
    return static_cast<uint32_t>(::crc32(crc, buf, static_cast<uInt>(len)));
}

}  // namespace mscrc
