// Single-query (decode step) GQA attention over the pre-allocated KV cache; replaces SDPA at q_len = 1
// ([O] Q2M:227-238 -> SDPA:40-104) and the torch.cat cache growth (CACHE:119-120: the cache here is written in
// place by rope_kv_append).  HBM-bound: every K and V byte of the live context is read exactly once per step:
// one CTA = one (sequence, KV head, 128-key chunk) and serves all G = H/Hkv query heads that share the KV head
// (scores: one thread per key, 16-byte loads of its K row, q broadcast from shared memory; P.V: one thread per output
// dim, coalesced V rows).  B*Hkv*ceil(Tmax/128) CTAs cover the GPU; a second kernel merges the chunk partials
// (log-sum-exp combine).  ctx_len lives in device memory so the launch parameters are step-invariant (CUDA-graph
// replay); chunks beyond the live context exit immediately.
// Algorithmic bytes per step: 2 (K,V) * Hkv * D * 2 B * ctx * B  (= 57344 B per token per sequence at 28 layers).
#include "common.h"
#include "ptx.cuh"

namespace af3 {

constexpr int DA_THREADS = 128;
constexpr int DA_CHUNK = 128;   // keys per CTA (one per thread in the score phase)
constexpr int DA_MAXG = 8;      // max query heads per KV head

template <int D>
__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                   float* __restrict__ part, int H, int Hkv, int Tmax, int nsplit, const int* __restrict__ ctx_len_p,
                   const int* __restrict__ kv_start, float scale) {
    static_assert(D == DA_THREADS, "one thread per output dim");
    const int G = H / Hkv;
    const int b = blockIdx.x, hk = blockIdx.y, sp = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int ctx = *ctx_len_p;
    const int start = kv_start ? kv_start[b] : 0;
    const int j0 = sp * DA_CHUNK;
    const int j_beg = max(j0, start), j_end = min(j0 + DA_CHUNK, ctx);
    float* pbase = part + ((static_cast<size_t>(b) * H + hk * G) * nsplit + sp) * (D + 2);
    if (j_beg >= j_end) {  // chunk entirely outside the live context: publish an empty partial
        for (int g = 0; g < G; ++g) {
            float* dst = pbase + static_cast<size_t>(g) * nsplit * (D + 2);
            if (tid == 0) {
                dst[D] = -INFINITY;
                dst[D + 1] = 0.f;
            }
        }
        return;
    }
    __shared__ __align__(16) float q_s[DA_MAXG][D];
    __shared__ float p_s[DA_MAXG][DA_CHUNK];
    __shared__ float red_m[DA_MAXG][4], red_l[DA_MAXG][4];

    const bf16* qrow = qkv + static_cast<size_t>(b) * (H + 2 * Hkv) * D + static_cast<size_t>(hk) * G * D;
    for (int i = tid; i < G * D; i += DA_THREADS) q_s[i / D][i % D] = __bfloat162float(qrow[i]) * scale;
    __syncthreads();

    // ---- scores: thread = key
    const int j = j0 + tid;
    const bool valid = j >= j_beg && j < j_end;
    float s[DA_MAXG];
#pragma unroll
    for (int g = 0; g < DA_MAXG; ++g) s[g] = 0.f;
    if (valid) {
        const uint4* kp = reinterpret_cast<const uint4*>(k_cache + ((static_cast<size_t>(b) * Hkv + hk) * Tmax + j) * D);
        uint4 kreg[D / 8];
#pragma unroll
        for (int c = 0; c < D / 8; ++c) kreg[c] = __ldg(kp + c);  // the whole 256-byte K row in flight at once
#pragma unroll
        for (int c = 0; c < D / 8; ++c) {
            const uint4 kv = kreg[c];
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
    }
    // ---- chunk softmax statistics per head
#pragma unroll
    for (int g = 0; g < DA_MAXG; ++g) {
        float mx = valid ? s[g] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red_m[g][warp] = mx;
    }
    __syncthreads();
    float m_c[DA_MAXG];
#pragma unroll
    for (int g = 0; g < DA_MAXG; ++g) {
        m_c[g] = fmaxf(fmaxf(red_m[g][0], red_m[g][1]), fmaxf(red_m[g][2], red_m[g][3]));
        const float p = valid ? __expf(s[g] - m_c[g]) : 0.f;
        p_s[g][tid] = bf16_round(p);  // P is rounded to bf16 before P.V as in flash-style SDPA kernels
        float ps = p;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
        if (lane == 0) red_l[g][warp] = ps;
    }
    __syncthreads();
    // ---- P.V: warp w takes keys jj = w, w+4, ...; lane owns dims 4*lane .. 4*lane+3 (8-byte loads, a warp reads a whole
    //      256-byte V row); 8 independent row loads in flight per warp; cross-warp reduction through shared memory
    float o_acc[DA_MAXG][4];
#pragma unroll
    for (int g = 0; g < DA_MAXG; ++g)
#pragma unroll
        for (int e = 0; e < 4; ++e) o_acc[g][e] = 0.f;
    const bf16* vbase = v_cache + ((static_cast<size_t>(b) * Hkv + hk) * Tmax + j0) * D + lane * 4;
    const int jj0 = j_beg - j0, jj1 = j_end - j0;
    for (int jb = jj0 + warp; jb < jj1; jb += 4 * 8) {
        uint2 vv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int jj = jb + 4 * u;
            vv[u] = (jj < jj1) ? __ldg(reinterpret_cast<const uint2*>(vbase + static_cast<size_t>(jj) * D)) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int jj = jb + 4 * u;
            if (jj < jj1) {
                const float2 va = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv[u].x));
                const float2 vb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&vv[u].y));
#pragma unroll
                for (int g = 0; g < DA_MAXG; ++g) {
                    if (g < G) {
                        const float p = p_s[g][jj];
                        o_acc[g][0] = fmaf(p, va.x, o_acc[g][0]);
                        o_acc[g][1] = fmaf(p, va.y, o_acc[g][1]);
                        o_acc[g][2] = fmaf(p, vb.x, o_acc[g][2]);
                        o_acc[g][3] = fmaf(p, vb.y, o_acc[g][3]);
                    }
                }
            }
        }
    }
    __syncthreads();  // q_s is dead from here: reuse it as the cross-warp reduction buffer [warp][G*D] (G*D <= 1024 floats)
    float* red = &q_s[0][0];
    // four passes (one per warp) keep the buffer at G*D floats: warp w adds its partial in turn
    for (int w = 0; w < 4; ++w) {
        if (warp == w) {
#pragma unroll
            for (int g = 0; g < DA_MAXG; ++g) {
                if (g < G) {
                    float4* dst = reinterpret_cast<float4*>(red + g * D + lane * 4);
                    float4 cur = (w == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : *dst;
                    cur.x += o_acc[g][0];
                    cur.y += o_acc[g][1];
                    cur.z += o_acc[g][2];
                    cur.w += o_acc[g][3];
                    *dst = cur;
                }
            }
        }
        __syncthreads();
    }
    for (int g = 0; g < G; ++g) {
        float* dst = pbase + static_cast<size_t>(g) * nsplit * (D + 2);
        dst[tid] = red[g * D + tid];
        if (tid == 0) {
            dst[D] = m_c[g];
            dst[D + 1] = red_l[g][0] + red_l[g][1] + red_l[g][2] + red_l[g][3];
        }
    }
}

template <int D>
__global__ void __launch_bounds__(D)
decode_attn_combine(const float* __restrict__ part, bf16* __restrict__ out, int H, int nsplit) {
    const int b = blockIdx.x, h = blockIdx.y, tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const float* src = part + (static_cast<size_t>(b) * H + h) * nsplit * (D + 2);
    float m = -INFINITY;
    for (int s = 0; s < nsplit; ++s) m = fmaxf(m, src[s * (D + 2) + D]);
    float l = 0.f, o = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float ms = src[s * (D + 2) + D];
        if (ms == -INFINITY) continue;  // empty chunk (its o slots were never written)
        const float w = __expf(ms - m);
        l += w * src[s * (D + 2) + D + 1];
        o += w * src[s * (D + 2) + tid];
    }
    out[(static_cast<size_t>(b) * H + h) * D + tid] = __float2bfloat16_rn(l > 0.f ? o / l : 0.f);
}

static int n_splits(int Tmax) { return ceil_div(Tmax, DA_CHUNK); }

size_t decode_attention_scratch_bytes(int B, int H, int D, int Tmax) {
    return static_cast<size_t>(B) * H * n_splits(Tmax) * (D + 2) * sizeof(float);
}

int decode_attention(cudaStream_t stream, const bf16* qkv, const bf16* k_cache, const bf16* v_cache, bf16* out,
                     float* scratch, int B, int H, int Hkv, int D, int Tmax, const int* ctx_len, const int* kv_start,
                     float scale) {
    AF3_REQUIRE(D == 128, "decode_attention: head_dim must be 128");
    AF3_REQUIRE(H % Hkv == 0 && H / Hkv <= DA_MAXG, "decode_attention: at most 8 query heads per KV head");
    AF3_REQUIRE(ctx_len != nullptr, "decode_attention: ctx_len must be a device pointer");
    const int ns = n_splits(Tmax);
    AF3_REQUIRE(ns <= 65535, "decode_attention: context too long");
    dim3 grid(B, Hkv, ns);
    AF3_CHECK_CUDA(launch_kernel(decode_attn_kernel<128>, grid, dim3(DA_THREADS), 0, stream, qkv, k_cache, v_cache, scratch, H,
                                 Hkv, Tmax, ns, ctx_len, kv_start, scale));
    dim3 g2(B, H);
    AF3_CHECK_CUDA(launch_kernel(decode_attn_combine<128>, g2, dim3(128), 0, stream, static_cast<const float*>(scratch), out, H, ns));
    return 0;
}

}  // namespace af3
