// Glue of the text memory bank around the LSTM recurrence, as three small kernels
// (ref: models/Multi_GCN_Multihead_att.py:366-398 — self.embedding(text); pack_padded_sequence; LSTM;
//  pad_packed_sequence(..., total_length): the bank is [B, L, 2H] with ZERO rows past each text's length).
//
// The recurrence works on compacted tokens (row offsets[b] + t of a [N, F] matrix), so the padded bank is
//   pad_rows_fwd   bank[b, t, :] = t < lens[b] ? y[offsets[b] + t, :] : 0        one pass, every bank row written once
//                  (torch: new_zeros of the whole bank + index_copy = 61 MB written twice, 97 us on the critical path
//                   of the training step at B = 512; this is 12 us)
//   pad_rows_bwd   gy[offsets[b] + t, :] = gbank[b, t, :] for t < lens[b]          (rows of gy beyond the valid tokens:
//                  zeroed by the caller)
// and the gradient of the embedding table is
//   embedding_bwd  gW[tok[i], :] += g[i, :]  (i over the compact rows, tok != padding_idx) with 128-bit vector
//                  reductions — one launch instead of ATen's sort / unique / segment-reduce chain (~15 launches, 0.3 ms as
//                  the tail of the step).  fp32 atomics: the sum order varies run to run, like the other atomics here.
#include "common.cuh"

namespace mgnns {

__global__ void __launch_bounds__(256) pad_rows_fwd_kernel(const float* __restrict__ y, const int32_t* __restrict__ offsets,
                                                           const int32_t* __restrict__ lens, int B, int L, int F4,
                                                           float* __restrict__ bank) {
    // one warp per bank row
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)B * L) return;
    const int b = (int)(row / L), t = (int)(row - (int64_t)b * L);
    const int lane = threadIdx.x & 31;
    float4* dst = reinterpret_cast<float4*>(bank) + row * F4;
    const int len = min(lens[b], L);
    if (t < len) {
        const float4* src = reinterpret_cast<const float4*>(y) + ((int64_t)offsets[b] + t) * F4;
        for (int f = lane; f < F4; f += 32) dst[f] = __ldg(src + f);
    } else {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int f = lane; f < F4; f += 32) dst[f] = z;
    }
}

__global__ void __launch_bounds__(256) pad_rows_bwd_kernel(const float* __restrict__ gbank, const int32_t* __restrict__ offsets,
                                                           const int32_t* __restrict__ lens, int B, int L, int F4,
                                                           float* __restrict__ gy) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= (int64_t)B * L) return;
    const int b = (int)(row / L), t = (int)(row - (int64_t)b * L);
    if (t >= min(lens[b], L)) return;
    const int lane = threadIdx.x & 31;
    const float4* src = reinterpret_cast<const float4*>(gbank) + row * F4;
    float4* dst = reinterpret_cast<float4*>(gy) + ((int64_t)offsets[b] + t) * F4;
    for (int f = lane; f < F4; f += 32) dst[f] = __ldg(src + f);
}

__global__ void __launch_bounds__(256) embedding_bwd_kernel(const int64_t* __restrict__ tok, const float* __restrict__ g,
                                                            int64_t n, int E4, int64_t padding_idx, int64_t V,
                                                            float* __restrict__ gw) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int64_t t = tok[row];
    if (t == padding_idx || t < 0 || t >= V) return;
    const int lane = threadIdx.x & 31;
    const float4* src = reinterpret_cast<const float4*>(g) + row * E4;
    float* dst = gw + t * (int64_t)E4 * 4;
    for (int f = lane; f < E4; f += 32) {
        const float4 v = __ldg(src + f);
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * f), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}

}  // namespace mgnns

using namespace mgnns;

extern "C" int mgnns_pad_rows_fwd(const float* y, const int32_t* offsets, const int32_t* lens, int B, int L, int F,
                                  float* bank, void* stream) {
    MG_REQUIRE(B >= 0 && L >= 1 && F >= 4 && F % 4 == 0, "pad_rows_fwd: F must be a positive multiple of 4");
    if (B == 0) return 0;
    MG_REQUIRE(y && offsets && lens && bank && aligned16(y) && aligned16(bank), "pad_rows_fwd: null or unaligned pointer");
    const int64_t rows = (int64_t)B * L;
    pad_rows_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(y, offsets, lens, B, L, F / 4, bank);
    MG_LAUNCH_CHECK("pad_rows_fwd");
    return 0;
}

extern "C" int mgnns_pad_rows_bwd(const float* gbank, const int32_t* offsets, const int32_t* lens, int B, int L, int F,
                                  float* gy, void* stream) {
    MG_REQUIRE(B >= 0 && L >= 1 && F >= 4 && F % 4 == 0, "pad_rows_bwd: F must be a positive multiple of 4");
    if (B == 0) return 0;
    MG_REQUIRE(gbank && offsets && lens && gy && aligned16(gbank) && aligned16(gy), "pad_rows_bwd: null or unaligned pointer");
    const int64_t rows = (int64_t)B * L;
    pad_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(gbank, offsets, lens, B, L, F / 4, gy);
    MG_LAUNCH_CHECK("pad_rows_bwd");
    return 0;
}

extern "C" int mgnns_embedding_bwd(const int64_t* tokens, const float* g, int64_t n, int E, int64_t padding_idx, int64_t V,
                                   float* gw, void* stream) {
    MG_REQUIRE(n >= 0 && E >= 4 && E % 4 == 0 && V >= 1, "embedding_bwd: E must be a positive multiple of 4");
    if (n == 0) return 0;
    MG_REQUIRE(tokens && g && gw && aligned16(g) && aligned16(gw), "embedding_bwd: null or unaligned pointer");
    embedding_bwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(tokens, g, n, E / 4, padding_idx, V, gw);
    MG_LAUNCH_CHECK("embedding_bwd");
    return 0;
}
