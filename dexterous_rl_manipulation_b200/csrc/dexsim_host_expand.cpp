// dexsim_host_expand.cpp -- host half of dexsim_step_host's default transport (plain C++, compiled by the host compiler).
//
// The five contact entries of the observation (envs/manipulation_env.py:262: `contacts.astype(float32)`, rows 40-44 of the
// [45, ld] observation) are 0/1 floats; dexsim_step_host downloads them as one byte per env (bit f = finger f) and this
// function writes the rows into the caller's pinned observation buffer while the DMA engine is still downloading the other
// rows.  It must keep ahead of that download (3 ms for 1 Mi envs): AVX2 with non-temporal stores when the CPU has it
// (0.3 ns per env), a scalar loop otherwise.
#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define DEXSIM_HAVE_X86 1
#endif

namespace dexsim {

static void expand_scalar(float* const rows[5], const uint8_t* mask, int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
        const unsigned b = mask[i];
        rows[0][i] = (float)(b & 1u); rows[1][i] = (float)((b >> 1) & 1u); rows[2][i] = (float)((b >> 2) & 1u);
        rows[3][i] = (float)((b >> 3) & 1u); rows[4][i] = (float)((b >> 4) & 1u);
    }
}

#ifdef DEXSIM_HAVE_X86
__attribute__((target("avx2"))) static void expand_avx2(float* const rows[5], const uint8_t* mask, int64_t lo, int64_t hi) {
    int64_t i = lo;
    for (; i < hi && (i & 7); ++i) expand_scalar(rows, mask, i, i + 1);          // up to a 32-byte boundary of the rows
    const __m256i one = _mm256_set1_epi32(1);
    for (; i + 8 <= hi; i += 8) {
        const __m256i v = _mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(mask + i)));
        _mm256_stream_ps(rows[0] + i, _mm256_cvtepi32_ps(_mm256_and_si256(v, one)));
        _mm256_stream_ps(rows[1] + i, _mm256_cvtepi32_ps(_mm256_and_si256(_mm256_srli_epi32(v, 1), one)));
        _mm256_stream_ps(rows[2] + i, _mm256_cvtepi32_ps(_mm256_and_si256(_mm256_srli_epi32(v, 2), one)));
        _mm256_stream_ps(rows[3] + i, _mm256_cvtepi32_ps(_mm256_and_si256(_mm256_srli_epi32(v, 3), one)));
        _mm256_stream_ps(rows[4] + i, _mm256_cvtepi32_ps(_mm256_and_si256(_mm256_srli_epi32(v, 4), one)));
    }
    _mm_sfence();
    if (i < hi) expand_scalar(rows, mask, i, hi);
}
#endif

// h_obs: [45, ld] float32 host buffer; mask: [ld] bytes; envs [lo, hi)
void expand_contact_rows_range(float* h_obs, const uint8_t* mask, int64_t lo, int64_t hi, int64_t ld) {
    float* rows[5];
    for (int f = 0; f < 5; ++f) rows[f] = h_obs + (size_t)(40 + f) * (size_t)ld;
#ifdef DEXSIM_HAVE_X86
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    // the streaming stores need 32-byte aligned rows: ld is a multiple of 32 floats by contract, the base may not be aligned
    if (have_avx2 && (reinterpret_cast<uintptr_t>(h_obs) & 31u) == 0 && (ld & 7) == 0) { expand_avx2(rows, mask, lo, hi); return; }
#endif
    expand_scalar(rows, mask, lo, hi);
}

}  // namespace dexsim
