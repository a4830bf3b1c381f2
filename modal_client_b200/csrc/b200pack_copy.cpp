// b200pack_copy.cpp -- the staging copy of the *_host entry points: pageable caller memory -> the pinned ring.
//
// The destination is written once and then read only by the GPU's copy engine, so the copy uses non-temporal
// stores: no read-for-ownership of the destination lines (2 instead of 3 memory transfers per byte) and the packer
// threads do not evict each other's source lines.  The widest store the CPU has is picked once at run time
// (AVX-512 64-byte stores where available: +10..14 % per thread over 32-byte stores on Sapphire Rapids, measured with
// tools/copybench.cpp; AVX2 otherwise; plain memcpy on anything older).  This file is host-only
// C++ (compiled by the host compiler, not by cudafe++), which is why it can carry per-function target attributes.
//
// Replaces nothing in the reference (it has no staging); it is the first stage of b200h_hash_batch_host /
// b200h_hash_files.  Exported as b200h_stream_copy (include/b200hash.h) so that callers filling page-locked memory
// themselves, and the CPU tests, use the very routine the library uses.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/b200hash.h"

namespace {

using copy_fn = void (*)(uint8_t*, const uint8_t*, size_t);

void copy_plain(uint8_t* dst, const uint8_t* src, size_t n) { memcpy(dst, src, n); }

__attribute__((target("avx2"))) void copy_avx2(uint8_t* dst, const uint8_t* src, size_t n) {
    if (n < 4096) {
        memcpy(dst, src, n);
        return;
    }
    const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;  // stores must be 32-byte aligned
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
        const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
    }
    _mm_sfence();  // the stores are weakly ordered: make them globally visible before the DMA is enqueued
    memcpy(dst + i, src + i, n - i);
}

__attribute__((target("avx512f"))) void copy_avx512(uint8_t* dst, const uint8_t* src, size_t n) {
    if (n < 4096) {
        memcpy(dst, src, n);
        return;
    }
    const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;  // stores must be 64-byte aligned
    memcpy(dst, src, head);
    dst += head; src += head; n -= head;
    size_t i = 0;
    for (; i + 256 <= n; i += 256) {
        const __m512i a = _mm512_loadu_si512(src + i);
        const __m512i b = _mm512_loadu_si512(src + i + 64);
        const __m512i c = _mm512_loadu_si512(src + i + 128);
        const __m512i d = _mm512_loadu_si512(src + i + 192);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i), a);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i + 64), b);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i + 128), c);
        _mm512_stream_si512(reinterpret_cast<__m512i*>(dst + i + 192), d);
    }
    _mm_sfence();
    memcpy(dst + i, src + i, n - i);
}

struct Choice {
    copy_fn fn;
    const char* name;
};

Choice choose() {
    const char* want = getenv("B200H_COPY_ISA");  // "avx512" | "avx2" | "plain": cap the choice (tests, experiments)
    const bool cap_avx2 = want && !strcmp(want, "avx2"), cap_plain = want && !strcmp(want, "plain");
    __builtin_cpu_init();
    if (!cap_plain && !cap_avx2 && __builtin_cpu_supports("avx512f")) return {copy_avx512, "avx512"};
    if (!cap_plain && __builtin_cpu_supports("avx2")) return {copy_avx2, "avx2"};
    return {copy_plain, "plain"};
}

const Choice& chosen() {
    static const Choice c = choose();  // thread-safe one-time initialisation
    return c;
}

}  // namespace

extern "C" void b200h_stream_copy(void* dst, const void* src, size_t n) {
    chosen().fn(static_cast<uint8_t*>(dst), static_cast<const uint8_t*>(src), n);
}

extern "C" const char* b200h_stream_copy_isa(void) { return chosen().name; }
