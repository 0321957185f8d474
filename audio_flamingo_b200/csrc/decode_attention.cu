// Single-query (decode step) GQA attention over the pre-allocated KV cache; replaces SDPA at q_len = 1
// ([O] Q2M:227-238 -> SDPA:40-104) and the torch.cat cache growth (CACHE:119-120: the cache here is written in
// place by rope_kv_append).  HBM-bound: every K and V byte of the live context is read exactly once per step.
//
// One CTA = one (sequence, KV head, 128-key chunk) and serves all G = H/Hkv (<= 16) query heads sharing the KV head.
// The problem is transposed so that the 128 keys / 128 output dims are the UMMA M dimension and the (padded) 16 query
// heads are N:
//     S^T[key, head] = K_tile[key, :] . Q[head, :]          tcgen05.mma 128x16x16 x 8,  A = K tile (K-major, via TMA)
//     O^T[dim, head] = V_tile^T[dim, key] . P^T[key, head]   tcgen05.mma 128x16x16 x 8,  A = V tile read MN-major
// so the K and V tiles stream HBM -> smem by TMA (2 x 32 KB in flight per CTA, 3 CTAs/SM) and never pass through
// registers; the softmax over the chunk is a cross-lane reduction of 16 columns (thread = key = TMEM lane).
// B*Hkv*ceil(Tmax/128) CTAs cover the GPU; the last-arriving chunk CTA of a (sequence, KV head) merges the chunk partials
// (log-sum-exp combine) -- no separate combine kernel.
// ctx_len lives in device memory so the launch parameters are step-invariant (CUDA-graph replay); chunks beyond the
// live context exit immediately.  Cache rows beyond the live context must hold finite values (the cache is
// zero-initialised): they are multiplied by P = 0.
// Algorithmic bytes per step: 2 (K,V) * Hkv * D * 2 B * ctx * B  (= 57344 B per token per sequence at 28 layers).
#include "common.h"
#include "ptx.cuh"

namespace af3 {

constexpr int DA_CHUNK = 128;    // keys per CTA
constexpr int DA_NH = 16;        // query heads per KV head, padded (UMMA N)
constexpr int DA_THREADS = 160;  // warps 0-3: softmax / epilogue (thread = TMEM lane), warp 4: TMA + MMA issue
constexpr int DA_D = 128;
constexpr int DA_SMEM = 2 * DA_CHUNK * DA_D * 2 /*K,V*/ + 2 * (DA_NH * 128) /*Q: 2 blocks of 16 x 128 B*/ +
                        2 * (DA_NH * 128) /*P*/ + 1024 /*align*/ + 1024 /*barriers + reduction scratch*/;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// Merge of the chunk partials of one (sequence, KV head), executed by the last-arriving chunk CTA.  Partial layout
// [b][h][split][D + 4] = (unnormalised o[D] relative to the chunk max, chunk max (log2 domain), chunk sum, pad).
// Work item = (head, 4 output dims): G * 32 items over the CTA's threads; the loads of all splits of an item are
// independent (two dependent rounds in total: maxima, then sums + outputs).  Chunks outside [start, ctx) never wrote
// theirs and are skipped by index.  Resets the arrival counter.
__device__ __forceinline__ void combine_heads(const float* __restrict__ part, bf16* __restrict__ out, int* counters, int b, int hk,
                                              int G, int H, int Hkv, int nsplit, int ctx, int start) {
    constexpr int D = DA_D, ST = D + 4, SB = 8;  // splits handled per unrolled block
    __threadfence();
    const int s_lo = start / DA_CHUNK, s_hi = (ctx + DA_CHUNK - 1) / DA_CHUNK;  // chunks that hold live keys
    for (int item = threadIdx.x; item < G * (D / 4); item += blockDim.x) {
        const int g = item / (D / 4), dq = item % (D / 4);
        const int h = hk * G + g;
        const float* src = part + (static_cast<size_t>(b) * H + h) * nsplit * ST;
        float m = -INFINITY;
        for (int s0 = s_lo; s0 < s_hi; s0 += SB) {
            float ms[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) ms[j] = (s0 + j < s_hi) ? __ldcg(src + (s0 + j) * ST + D) : -INFINITY;
#pragma unroll
            for (int j = 0; j < SB; ++j) m = fmaxf(m, ms[j]);
        }
        float l = 0.f;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = s_lo; s0 < s_hi; s0 += SB) {
            float2 ml[SB];
            float4 ov[SB];
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                const bool ok = s0 + j < s_hi;
                ml[j] = ok ? __ldcg(reinterpret_cast<const float2*>(src + (s0 + j) * ST + D)) : make_float2(-INFINITY, 0.f);
                ov[j] = ok ? __ldcg(reinterpret_cast<const float4*>(src + (s0 + j) * ST + dq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < SB; ++j) {
                const float w = (ml[j].x == -INFINITY) ? 0.f : exp2f(ml[j].x - m);
                l += w * ml[j].y;
                o.x += w * ov[j].x;
                o.y += w * ov[j].y;
                o.z += w * ov[j].z;
                o.w += w * ov[j].w;
            }
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        bf16* dst = out + (static_cast<size_t>(b) * H + h) * D + dq * 4;
        *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16x2(o.x * inv, o.y * inv), pack_bf16x2(o.z * inv, o.w * inv));
    }
    if (threadIdx.x == 0) counters[b * Hkv + hk] = 0;
}

__global__ void __launch_bounds__(DA_THREADS)
decode_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, float* __restrict__ part, bf16* __restrict__ out,
                   int* __restrict__ counters, int H, int Hkv, int nsplit, const int* __restrict__ ctx_len_p,
                   const int* __restrict__ kv_start, float scale_log2) {
    constexpr int D = DA_D;
    const int G = H / Hkv;
    const int b = blockIdx.x, hk = blockIdx.y, sp = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const int ctx = *ctx_len_p;
    const int start = kv_start ? kv_start[b] : 0;
    const int j0 = sp * DA_CHUNK;
    const int j_beg = max(j0, start), j_end = min(j0 + DA_CHUNK, ctx);
    float* pbase = part + ((static_cast<size_t>(b) * H + hk * G) * nsplit + sp) * (D + 4);
    if (j_beg >= j_end) {  // chunk entirely outside the live context: it only counts as arrived
        __shared__ int last_flag;
        if (tid == 0) last_flag = (atomicAdd(counters + b * Hkv + hk, 1) == nsplit - 1);
        __syncthreads();
        if (last_flag) combine_heads(part, out, counters, b, hk, G, H, Hkv, nsplit, ctx, start);
        return;
    }
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sK = smem;                   // 2 blocks [128 keys x 128 B]
    uint8_t* sV = sK + DA_CHUNK * D * 2;  // 2 blocks [128 keys x 128 B]
    uint8_t* sQ = sV + DA_CHUNK * D * 2;  // 2 blocks [16 heads x 128 B]
    uint8_t* sP = sQ + 2 * DA_NH * 128;   // 2 blocks [16 heads x 128 B]  (64 keys along each 128-byte row)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * DA_NH * 128);
    uint64_t *qk_full = bars, *v_full = bars + 1, *s_full = bars + 2, *p_full = bars + 3, *o_full = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
    float* red = reinterpret_cast<float*>(bars + 6);  // [2][4 warps][16 heads]

    if (tid == 128) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_init(qk_full, 1);
        mbar_init(v_full, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_slot, 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_S = *tmem_slot, tmem_O = tmem_S + DA_NH;

    if (warp == 4) {
        if (lane == 0) {
            // all three operands in flight at once
            mbar_arrive_expect_tx(qk_full, DA_CHUNK * D * 2 + 2 * DA_NH * 128);
#pragma unroll
            for (int db = 0; db < 2; ++db) {
                tma_load_2d(sQ + db * DA_NH * 128, &map_q, qk_full, db * 64, b * (H + 2 * Hkv) + hk * G);
                tma_load_3d(sK + db * 16384, &map_k, qk_full, db * 64, j0, b * Hkv + hk);
            }
            mbar_arrive_expect_tx(v_full, DA_CHUNK * D * 2);
#pragma unroll
            for (int db = 0; db < 2; ++db) tma_load_3d(sV + db * 16384, &map_v, v_full, db * 64, j0, b * Hkv + hk);
            constexpr uint32_t idesc_s = make_idesc_bf16(128, DA_NH, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(128, DA_NH, 1, 0);  // A (= V tile) is MN-major
            const uint32_t aK = smem_u32(sK), aQ = smem_u32(sQ), aV = smem_u32(sV), aP = smem_u32(sP);
            // S^T = K . Q^T
            mbar_wait(qk_full, 0);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < D / 16; ++kk)
                umma_bf16_ss(tmem_S, make_smem_desc_sw128(aK + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                             make_smem_desc_sw128(aQ + (kk >> 2) * (DA_NH * 128) + (kk & 3) * 32, 0, 1024), idesc_s, kk != 0);
            umma_commit(s_full);
            // O^T = V^T . P^T
            mbar_wait(p_full, 0);
            mbar_wait(v_full, 0);
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < DA_CHUNK / 16; ++kk)
                umma_bf16_ss(tmem_O, make_smem_desc_sw128(aV + kk * 2048, 16384, 1024),
                             make_smem_desc_sw128(aP + (kk >> 2) * (DA_NH * 128) + (kk & 3) * 32, 0, 1024), idesc_o, kk != 0);
            umma_commit(o_full);
        }
        __syncwarp();
    } else {
        // ---- softmax over the chunk: thread = key (TMEM lane), 16 columns = heads
        const int j = j0 + tid;
        const bool valid = j >= j_beg && j < j_end;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        mbar_wait(s_full, 0);
        tc_fence_after();
        uint32_t sv[16];
        tmem_ld16(tmem_S + lane_off, sv);
        tmem_ld_wait();
        float t[DA_NH], m[DA_NH], p[DA_NH];
#pragma unroll
        for (int g = 0; g < DA_NH; ++g) {
            t[g] = (valid && g < G) ? __uint_as_float(sv[g]) * scale_log2 : -INFINITY;
            float mx = t[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) red[warp * DA_NH + g] = mx;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int g = 0; g < DA_NH; ++g) {
            m[g] = fmaxf(fmaxf(red[g], red[DA_NH + g]), fmaxf(red[2 * DA_NH + g], red[3 * DA_NH + g]));
            p[g] = (t[g] == -INFINITY) ? 0.f : exp2f(t[g] - m[g]);
            float ps = p[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
            if (lane == 0) red[64 + warp * DA_NH + g] = ps;
        }
        // P^T[key = tid][head g] -> B tile [16 heads][128 keys], K-major, 128B swizzle: row g, key column tid
        {
            uint8_t* blk = sP + (tid >> 6) * (DA_NH * 128);
            const int kc = tid & 63;
#pragma unroll
            for (int g = 0; g < DA_NH; ++g)
                *reinterpret_cast<bf16*>(blk + g * 128 + ((((kc >> 3) ^ (g & 7)) << 4) | ((kc & 7) << 1))) = __float2bfloat16_rn(p[g]);
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
        asm volatile("bar.sync 1, 128;" ::: "memory");  // red[64..] complete
        // ---- O^T[dim = tid][head] -> partial
        mbar_wait(o_full, 0);
        tc_fence_after();
        uint32_t ov[16];
        tmem_ld16(tmem_O + lane_off, ov);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < DA_NH; ++g) {
            if (g < G) {
                float* dst = pbase + static_cast<size_t>(g) * nsplit * (D + 4);
                dst[tid] = __uint_as_float(ov[g]);
                if (tid == 0) {
                    dst[D] = m[g];
                    dst[D + 1] = red[64 + g] + red[64 + DA_NH + g] + red[64 + 2 * DA_NH + g] + red[64 + 3 * DA_NH + g];
                }
            }
        }
    }
    tc_fence_before();
    __threadfence();  // partials visible before this chunk is counted
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_S, 32);
    // ---- the last-arriving chunk CTA of this (sequence, KV head) merges all chunk partials (log-sum-exp combine):
    //      replaces a separate combine kernel per layer
    int* flag = reinterpret_cast<int*>(red);
    if (tid == 0) *flag = (atomicAdd(counters + b * Hkv + hk, 1) == nsplit - 1);
    __syncthreads();
    if (*flag) combine_heads(part, out, counters, b, hk, G, H, Hkv, nsplit, ctx, start);
}

static int n_splits(int Tmax) { return ceil_div(Tmax, DA_CHUNK); }

// scratch = chunk partials + one arrival counter per (sequence, KV head); the counters must be ZERO on first use (the kernel
// leaves them zero)
static size_t partial_bytes(int B, int H, int D, int Tmax) {
    return (static_cast<size_t>(B) * H * n_splits(Tmax) * (D + 4) * sizeof(float) + 255) & ~static_cast<size_t>(255);
}
size_t decode_attention_scratch_bytes(int B, int H, int D, int Tmax) {
    return partial_bytes(B, H, D, Tmax) + static_cast<size_t>(B) * H * sizeof(int);
}

int decode_attention(cudaStream_t stream, const bf16* qkv, const bf16* k_cache, const bf16* v_cache, bf16* out,
                     float* scratch, int B, int H, int Hkv, int D, int Tmax, const int* ctx_len, const int* kv_start,
                     float scale) {
    AF3_REQUIRE(D == 128, "decode_attention: head_dim must be 128");
    AF3_REQUIRE(H % Hkv == 0 && H / Hkv <= DA_NH, "decode_attention: at most 16 query heads per KV head");
    AF3_REQUIRE(ctx_len != nullptr, "decode_attention: ctx_len must be a device pointer");
    const int ns = n_splits(Tmax);
    AF3_REQUIRE(ns <= 65535, "decode_attention: context too long");
    static bool configured = false;
    if (!configured) {
        AF3_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DA_SMEM));
        configured = true;
    }
    CUtensorMap mq, mk, mv;
    // q heads as rows of 128: the packed projection row of sequence b holds (H + 2 Hkv) such rows
    if (int e = make_tmap_2d(&mq, qkv, D, static_cast<uint64_t>(B) * (H + 2 * Hkv), D, 64, DA_NH)) return e;
    if (int e = make_tmap_3d(&mk, k_cache, D, Tmax, static_cast<uint64_t>(B) * Hkv, D, static_cast<uint64_t>(Tmax) * D, 64,
                             DA_CHUNK, 1))
        return e;
    if (int e = make_tmap_3d(&mv, v_cache, D, Tmax, static_cast<uint64_t>(B) * Hkv, D, static_cast<uint64_t>(Tmax) * D, 64,
                             DA_CHUNK, 1))
        return e;
    dim3 grid(B, Hkv, ns);
    int* counters = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(scratch) + partial_bytes(B, H, D, Tmax));
    AF3_CHECK_CUDA(launch_kernel(decode_attn_kernel, grid, dim3(DA_THREADS), DA_SMEM, stream, mq, mk, mv, scratch, out, counters, H,
                                 Hkv, ns, ctx_len, kv_start, scale * 1.4426950408889634f));
    return 0;
}

}  // namespace af3
