// Single-query (decode step) GQA attention over the pre-allocated KV cache; replaces SDPA at q_len = 1
// ([O] Q2M:227-238 -> SDPA:40-104) and the torch.cat cache growth (CACHE:119-120: the cache here is written in
// place by rope_kv_append).  HBM-bound: every K and V byte of the live context is read exactly once per step
// (one CTA serves all G = H/Hkv query heads that share a KV head), 16-byte loads, split over the context so that
// B*Hkv*nsplit CTAs cover the 148 SMs; a second kernel merges the split partials (log-sum-exp combine).
// ctx_len lives in device memory so the launch parameters are step-invariant (CUDA-graph replay).
// Algorithmic bytes per step: 2 (K,V) * Hkv * D * 2 B * ctx * B  (= 57344 B per token per sequence at 28 layers).
#include "common.h"
#include "ptx.cuh"

namespace af3 {

constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 128;   // keys per inner chunk (one per thread in the score phase)
constexpr int DA_MAXG = 8;      // max query heads per KV head
constexpr int DA_NSPLIT = 4;

template <int D>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                   float* __restrict__ part, int H, int Hkv, int Tmax, const int* __restrict__ ctx_len_p,
                   const int* __restrict__ kv_start, float scale) {
    static_assert(D == DA_THREADS, "one thread per output dim");
    const int G = H / Hkv;
    const int b = blockIdx.x, hk = blockIdx.y, sp = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ctx = *ctx_len_p;
    const int start = kv_start ? kv_start[b] : 0;
    __shared__ __align__(16) float q_s[DA_MAXG][D];
    __shared__ float p_s[DA_MAXG][DA_CHUNK];
    __shared__ float red[DA_MAXG][4];

    const bf16* qrow = qkv + static_cast<size_t>(b) * (H + 2 * Hkv) * D + static_cast<size_t>(hk) * G * D;
    for (int i = tid; i < G * D; i += DA_THREADS) q_s[i / D][i % D] = __bfloat162float(qrow[i]) * scale;
    __syncthreads();

    // this split's key range
    const int span = max(ctx - start, 0);
    const int per = (span + DA_NSPLIT - 1) / DA_NSPLIT;
    const int j_beg = start + sp * per;
    const int j_end = min(start + (sp + 1) * per, ctx);

    const bf16* kbase = k_cache + (static_cast<size_t>(b) * Hkv + hk) * Tmax * D;
    const bf16* vbase = v_cache + (static_cast<size_t>(b) * Hkv + hk) * Tmax * D;

    float m_run[DA_MAXG], l_run[DA_MAXG], o_run[DA_MAXG];
#pragma unroll
    for (int g = 0; g < DA_MAXG; ++g) {
        m_run[g] = -INFINITY;
        l_run[g] = 0.f;
        o_run[g] = 0.f;
    }

    for (int j0 = j_beg; j0 < j_end; j0 += DA_CHUNK) {
        const int j = j0 + tid;
        float s[DA_MAXG];
#pragma unroll
        for (int g = 0; g < DA_MAXG; ++g) s[g] = 0.f;
        if (j < j_end) {
            const uint4* kp = reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(j) * D);
#pragma unroll 4
            for (int c = 0; c < D / 8; ++c) {
                const uint4 kv = __ldg(kp + c);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&kv);
                float kf[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 t = __bfloat1622float2(h2[e]);
                    kf[2 * e] = t.x;
                    kf[2 * e + 1] = t.y;
                }
#pragma unroll
                for (int g = 0; g < DA_MAXG; ++g) {
                    if (g < G) {
                        const float4 qa = *reinterpret_cast<const float4*>(&q_s[g][c * 8]);
                        const float4 qb = *reinterpret_cast<const float4*>(&q_s[g][c * 8 + 4]);
                        s[g] += kf[0] * qa.x + kf[1] * qa.y + kf[2] * qa.z + kf[3] * qa.w + kf[4] * qb.x + kf[5] * qb.y +
                                kf[6] * qb.z + kf[7] * qb.w;
                    }
                }
            }
        } else {
#pragma unroll
            for (int g = 0; g < DA_MAXG; ++g) s[g] = -INFINITY;
        }
        // chunk max per head (block reduce), update running max, write probabilities
        float alpha[DA_MAXG];
#pragma unroll
        for (int g = 0; g < DA_MAXG; ++g) {
            float mx = s[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) red[g][warp] = mx;
        }
        __syncthreads();
#pragma unroll
        for (int g = 0; g < DA_MAXG; ++g) {
            const float cm = fmaxf(fmaxf(red[g][0], red[g][1]), fmaxf(red[g][2], red[g][3]));
            const float mn = fmaxf(m_run[g], cm);
            alpha[g] = (m_run[g] == -INFINITY) ? 0.f : __expf(m_run[g] - mn);
            m_run[g] = mn;
            const float p = (s[g] == -INFINITY) ? 0.f : __expf(s[g] - mn);
            p_s[g][tid] = bf16_round(p);  // P is rounded to bf16 before P.V as in flash-style SDPA kernels
            float ps = p;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
            __syncthreads();  // red[g] reads done before reuse; also orders p_s writes
            if (lane == 0) red[g][warp] = ps;
            __syncthreads();
            l_run[g] = l_run[g] * alpha[g] + (red[g][0] + red[g][1] + red[g][2] + red[g][3]);
            o_run[g] *= alpha[g];
        }
        __syncthreads();
        // P.V: thread = output dim
        const int nk = min(DA_CHUNK, j_end - j0);
        for (int jj = 0; jj < nk; ++jj) {
            const float v = __bfloat162float(vbase[static_cast<size_t>(j0 + jj) * D + tid]);
#pragma unroll
            for (int g = 0; g < DA_MAXG; ++g)
                if (g < G) o_run[g] = fmaf(p_s[g][jj], v, o_run[g]);
        }
        __syncthreads();
    }
    // partials: [b][h][sp][D + 2]
    for (int g = 0; g < G; ++g) {
        float* dst = part + ((static_cast<size_t>(b) * H + hk * G + g) * DA_NSPLIT + sp) * (D + 2);
        dst[tid] = o_run[g];
        if (tid == 0) {
            dst[D] = m_run[g];
            dst[D + 1] = l_run[g];
        }
    }
}

template <int D>
__global__ void __launch_bounds__(D)
decode_attn_combine(const float* __restrict__ part, bf16* __restrict__ out, int H) {
    const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    const float* src = part + (static_cast<size_t>(b) * H + h) * DA_NSPLIT * (D + 2);
    float m = -INFINITY;
#pragma unroll
    for (int s = 0; s < DA_NSPLIT; ++s) m = fmaxf(m, src[s * (D + 2) + D]);
    float l = 0.f, o = 0.f;
#pragma unroll
    for (int s = 0; s < DA_NSPLIT; ++s) {
        const float ms = src[s * (D + 2) + D];
        const float w = (ms == -INFINITY) ? 0.f : __expf(ms - m);
        l += w * src[s * (D + 2) + D + 1];
        o += w * src[s * (D + 2) + tid];
    }
    out[(static_cast<size_t>(b) * H + h) * D + tid] = __float2bfloat16_rn(l > 0.f ? o / l : 0.f);
}

size_t decode_attention_scratch_bytes(int B, int H, int D) {
    return static_cast<size_t>(B) * H * DA_NSPLIT * (D + 2) * sizeof(float);
}

int decode_attention(cudaStream_t stream, const bf16* qkv, const bf16* k_cache, const bf16* v_cache, bf16* out,
                     float* scratch, int B, int H, int Hkv, int D, int Tmax, const int* ctx_len, const int* kv_start,
                     float scale) {
    AF3_REQUIRE(D == 128, "decode_attention: head_dim must be 128");
    AF3_REQUIRE(H % Hkv == 0 && H / Hkv <= DA_MAXG, "decode_attention: at most 8 query heads per KV head");
    AF3_REQUIRE(ctx_len != nullptr, "decode_attention: ctx_len must be a device pointer");
    dim3 grid(B, Hkv, DA_NSPLIT);
    decode_attn_kernel<128><<<grid, DA_THREADS, 0, stream>>>(qkv, k_cache, v_cache, scratch, H, Hkv, Tmax, ctx_len,
                                                            kv_start, scale);
    AF3_CHECK_LAUNCH();
    dim3 g2(B, H);
    decode_attn_combine<128><<<g2, 128, 0, stream>>>(scratch, out, H);
    AF3_CHECK_LAUNCH();
    return 0;
}

}  // namespace af3
