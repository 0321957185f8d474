// Fused softmax attention, second generation (round 2), for the same two shapes as attention_tcgen05.cu:
//   * AF-Whisper encoder self-attention: bidirectional, 20 heads x 64, 1500 frames, key-padding mask   ([O] AF3M:116-189)
//   * Qwen2 decoder prefill: causal GQA 28:4 x 128 with left padding / live-cache offset                 ([O] Q2M:206-245)
// replacing F.scaled_dot_product_attention ([O] SDPA:40-104).
//
// What round 1's profile said (profiles/r01i_ncu_attention_source.md): tensor pipe 20 % active; one 128-row Q tile per CTA made
// QK^T -> softmax -> PV a serial chain, the softmax warps spent 30 % of their samples polling for S and 12 % draining the
// P stores to shared memory before the proxy fence, and two threads per row needed a block barrier per tile to exchange maxima.
// This kernel changes the structure:
//   * one CTA = TWO 128-row Q tiles (A, B) of one (batch, head) that share every K/V tile (loaded once, used twice) and
//     ping-pong: while softmax group A works on S_A the tensor core runs group B's MMAs and vice versa;
//   * one thread per query row (TMEM lane): row max and row sum are thread-local -- no shuffles, no barrier, no smem;
//   * P never leaves tensor memory: the softmax thread overwrites its own S row in place with bf16 P (tcgen05.st, two keys
//     per 32-bit cell) and O += P V reads A = P straight from TMEM (tcgen05.mma with a TMEM A operand); no P buffer in
//     shared memory, no fence.proxy.async;
//   * O stays in TMEM and is rescaled lazily (only when the running max grew by > 2^8), as before.
// Roles (10 warps): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 softmax group A, warps 6-9 softmax group B.
// TMEM (512 columns): S_A | S_B (128 each, P aliased onto the first 64) | O_A | O_B (D each).
// Issue order of the MMA thread in steady state:  PV_A(j)  QK_A(j+1)  PV_B(j)  QK_B(j+1)  -- tcgen05.mma executes in issue
// order, which is what makes the in-place S -> P -> S reuse safe: QK_g(j+1) is issued after PV_g(j) has consumed P_g(j).
#include <stdlib.h>

#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace af3 {

constexpr int A2_BM = 128;  // rows per Q tile (two per CTA)
constexpr int A2_BN = 128;  // keys per K/V tile
constexpr float A2_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float a2_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Attn2Args {
    int Tq, Tk, H, Hkv, causal, kv_layout;
    float scale_log2;
    const int* kv_len;
    const int* kv_start;
    bf16* out;
    int ldo;
};

template <int D>
struct Attn2Cfg {
    static constexpr int NSTG = (D == 128) ? 4 : 6;         // K/V ring slots (one K or V tile each)
    static constexpr int TILE_BYTES = A2_BN * D * 2;
    static constexpr int Q_BYTES = 2 * A2_BM * D * 2;       // both Q tiles
    static constexpr int SMEM_BYTES = Q_BYTES + NSTG * TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 6144 /*half-row max / sum exchange (TPR = 2)*/;
    static constexpr int TMEM_COLS = 512;
};

// key tiles [j_lo, j_hi) a 128-row Q tile starting at q0 has to visit
__device__ __forceinline__ void a2_tile_range(const Attn2Args& a, int b, int q0, int& j_lo, int& j_hi) {
    const int kvl = a.kv_len ? min(a.kv_len[b], a.Tk) : a.Tk;
    const int kvs = a.kv_start ? a.kv_start[b] : 0;
    int hi_key = kvl;  // exclusive
    if (a.causal) hi_key = min(hi_key, q0 + A2_BM - 1 + (a.Tk - a.Tq) + 1);
    j_lo = kvs / A2_BN;
    j_hi = (hi_key + A2_BN - 1) / A2_BN;
    if (j_hi < j_lo) j_hi = j_lo;
}

// TPR = threads per query row in the softmax groups.  1: 4 warps per group, whole row per thread (no exchange at all).
// 2: 8 warps per group, each thread owns 64 of the 128 key columns, HOLDS them in registers (one TMEM pass instead of two) and the
// two half-row maxima / sums are exchanged through shared memory with one 256-thread named barrier per tile.  The ncu profile of
// TPR = 1 (profiles/r02c_ncu_attention_v2_d128.md) showed the softmax to be latency-bound per warp (1260 instructions per tile at
// 0.18 IPC, two such warps per scheduler): TPR = 2 halves the per-thread chain and doubles the warps each scheduler can pick from.
// Measured (profiles/r02i_microbench_attention.json): TPR = 2 is 7-12 % SLOWER than TPR = 1 on all three shapes, so TPR = 1 is the default.
template <int D, int TPR>
__global__ void __launch_bounds__(64 + 256 * TPR, 1)
attention2_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                  const __grid_constant__ CUtensorMap map_v, const Attn2Args a) {
    using Cfg = Attn2Cfg<D>;
    constexpr int DB = D / 64;          // 64-wide (128-byte) column blocks per tile
    constexpr int NSTG = Cfg::NSTG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sQ = smem;                                 // tile A then tile B, each DB blocks of [128 rows x 128 B]
    uint8_t* sKV = sQ + Cfg::Q_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + NSTG * Cfg::TILE_BYTES);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + NSTG;
    uint64_t* s_full = kv_empty + NSTG;   // [2]
    uint64_t* p_full = s_full + 2;        // [2]
    uint64_t* o_full = p_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);
    [[maybe_unused]] float* xch = reinterpret_cast<float*>(bars + 32);   // [2 groups][3: max parity 0 / max parity 1 / sums][2 halves][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // heavy (late) causal row blocks first: the last wave is then made of the short ones
    const int qblk = a.causal ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
    const int q0 = qblk * 2 * A2_BM, h = blockIdx.y, b = blockIdx.z;
    const int hk = h / (a.H / a.Hkv);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < NSTG; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1);
            mbar_init(&p_full[g], 128 * TPR);
            mbar_init(&o_full[g], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // S_g at columns [g*128, +128) (P_g aliased onto its first 64), O_g at 256 + g*D
    const uint32_t tmem_S0 = tmem_base, tmem_O0 = tmem_base + 2 * A2_BN;

    // tile ranges of the two Q tiles (B lies 128 rows below A: under a causal mask it sees at least as many key tiles)
    int j_lo, j_hiA, j_hiB, j_lo_b;
    a2_tile_range(a, b, q0, j_lo, j_hiA);
    a2_tile_range(a, b, q0 + A2_BM, j_lo_b, j_hiB);
    const bool activeB = q0 + A2_BM < a.Tq;
    const int nA = j_hiA - j_lo;
    const int nB = activeB ? (j_hiB - j_lo) : 0;
    const int n = max(nA, nB);

    if (warp == 0) {
        if (lane == 0 && n > 0) {
            mbar_arrive_expect_tx(q_full, activeB ? Cfg::Q_BYTES : Cfg::Q_BYTES / 2);
#pragma unroll
            for (int db = 0; db < DB; ++db) tma_load_3d(sQ + db * 16384, &map_q, q_full, h * D + db * 64, q0, b);
            if (activeB) {
#pragma unroll
                for (int db = 0; db < DB; ++db)
                    tma_load_3d(sQ + Cfg::Q_BYTES / 2 + db * 16384, &map_q, q_full, h * D + db * 64, q0 + A2_BM, b);
            }
            int slot = 0;
            uint32_t phase = 0;
            for (int j = j_lo; j < j_lo + n; ++j) {
#pragma unroll
                for (int kv = 0; kv < 2; ++kv) {   // load 2t = K_t, load 2t+1 = V_t
                    mbar_wait(&kv_empty[slot], phase ^ 1);
                    mbar_arrive_expect_tx(&kv_full[slot], Cfg::TILE_BYTES);
                    uint8_t* dst = sKV + slot * Cfg::TILE_BYTES;
                    const CUtensorMap* mp = kv ? &map_v : &map_k;
#pragma unroll
                    for (int db = 0; db < DB; ++db) {
                        if (a.kv_layout)
                            tma_load_3d(dst + db * 16384, mp, &kv_full[slot], db * 64, j * A2_BN, b * a.Hkv + hk);
                        else
                            tma_load_3d(dst + db * 16384, mp, &kv_full[slot], hk * D + db * 64, j * A2_BN, b);
                    }
                    if (++slot == NSTG) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && n > 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(A2_BM, A2_BN, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_bf16(A2_BM, D, 0, 1);  // A = P from TMEM (K-major), B = V MN-major from the row-major tile
            const uint32_t aQ = smem_u32(sQ);
            mbar_wait(q_full, 0);
            auto wait_load = [&](int i) {   // i-th load of the producer's sequence; returns its smem address
                const int slot = i % NSTG;
                mbar_wait(&kv_full[slot], (i / NSTG) & 1);
                tc_fence_after();
                return smem_u32(sKV + slot * Cfg::TILE_BYTES);
            };
            auto release_load = [&](int i) { umma_commit(&kv_empty[i % NSTG]); };
            auto qk = [&](int g, int t) {   // S_g = Q_g K_t^T
                const uint32_t aK = wait_load(2 * t);
                const uint32_t aQg = aQ + g * (Cfg::Q_BYTES / 2);
                const uint32_t dS = tmem_S0 + g * A2_BN;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
                    umma_bf16_ss(dS, make_smem_desc_sw128(aQg + off, 0, 1024), make_smem_desc_sw128(aK + off, 0, 1024), idesc_s, kk != 0);
                }
                umma_commit(&s_full[g]);
            };
            auto pv = [&](int g, int t) {   // O_g += P_g V_t, P_g read from TMEM
                mbar_wait(&p_full[g], t & 1);
                const uint32_t aV = wait_load(2 * t + 1);
                const uint32_t dO = tmem_O0 + g * D, aP = tmem_S0 + g * A2_BN;
#pragma unroll
                for (int kk = 0; kk < A2_BN / 16; ++kk)
                    umma_bf16_ts(dO, aP + kk * 8, make_smem_desc_sw128(aV + kk * 2048, 16384, 1024), idesc_o, (t | kk) != 0);
                // o_full only announces the LAST product (the epilogue waits for it).  In between, the softmax group knows that
                // P_g(t-1) V has retired as soon as it has seen s_full for tile t: Q K_t^T was issued after it and tcgen05.mma
                // retires in issue order.  (Committing o_full every tile completed mbarrier phases nobody waited for, which
                // compute-sanitizer's synccheck reports as an error.)
                if (t + 1 == (g ? nB : nA)) umma_commit(&o_full[g]);
            };
            if (nA > 0) qk(0, 0);
            if (nB > 0) qk(1, 0);
            release_load(0);
            for (int t = 0; t < n; ++t) {
                if (t < nA) pv(0, t);
                if (t + 1 < nA) qk(0, t + 1);
                if (t < nB) pv(1, t);
                release_load(2 * t + 1);                 // V_t: both PVs issued
                if (t + 1 < nB) qk(1, t + 1);
                if (t + 1 < n) release_load(2 * t + 2);  // K_{t+1}: both QKs issued
            }
        }
        __syncwarp();
    } else {
        if constexpr (TPR == 2) {
            // ---- softmax, two threads per query row: warp w of the group handles TMEM lane quarter (w & 3) and key columns
            //      [half * 64, half * 64 + 64) with half = (w - first warp of the group) >> 2
            const int sw = warp - 2;
            const int g = sw >> 3;
            const int half = (sw >> 2) & 1;
            const int qd = warp & 3;
            const int row = qd * 32 + lane;
            const int q0g = q0 + g * A2_BM;
            const int qi = q0g + row;
            const int n_g = g ? nB : nA;
            const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
            const uint32_t s_col = tmem_S0 + g * A2_BN + lane_off;
            constexpr int OC = D / 2;               // O columns owned by this thread (rescale / epilogue)
            const uint32_t o_col = tmem_O0 + g * D + lane_off + half * OC;
            const int kvl = a.kv_len ? min(a.kv_len[b], a.Tk) : a.Tk;
            const int kvs = a.kv_start ? a.kv_start[b] : 0;
            const int causal_hi = a.causal ? qi + (a.Tk - a.Tq) : 0x7fffffff;
            const float sl2 = a.scale_log2;
            const int cbase = half * 64;
            float* xg = xch + g * 768;              // this group's exchange area: [3][2][128]
            float m_ref = -INFINITY, l_run = 0.f;   // l_run: partial row sum over this thread's columns

            for (int t = 0; t < n_g; ++t) {
                const int k0 = (j_lo + t) * A2_BN;
                mbar_wait(&s_full[g], t & 1);
                tc_fence_after();
                const bool need_mask = (k0 < kvs) || (k0 + A2_BN > kvl) || (a.causal && k0 + A2_BN - 1 > q0g + (a.Tk - a.Tq));
                const int vlo = max(kvs - k0, 0);
                const int vhi = min(min(kvl - 1, causal_hi) - k0, A2_BN - 1);
                const uint32_t vspan = static_cast<uint32_t>(vhi - vlo);
                const bool row_empty = vhi < vlo;
                // the thread's 64 scores: ONE TMEM read, kept in registers for both the max and the exponentials
                uint32_t s0[32], s1[32];
                tmem_ld32(s_col + cbase, s0);
                tmem_ld32(s_col + cbase + 32, s1);
                tmem_ld_wait();
                float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                if (need_mask) {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        const bool ok0 = static_cast<uint32_t>(cbase + e - vlo) <= vspan;
                        const bool ok1 = static_cast<uint32_t>(cbase + e + 1 - vlo) <= vspan;
                        const bool ok2 = static_cast<uint32_t>(cbase + 32 + e - vlo) <= vspan;
                        const bool ok3 = static_cast<uint32_t>(cbase + 33 + e - vlo) <= vspan;
                        mx0 = fmaxf(mx0, ok0 ? __uint_as_float(s0[e]) : -INFINITY);
                        mx1 = fmaxf(mx1, ok1 ? __uint_as_float(s0[e + 1]) : -INFINITY);
                        mx2 = fmaxf(mx2, ok2 ? __uint_as_float(s1[e]) : -INFINITY);
                        mx3 = fmaxf(mx3, ok3 ? __uint_as_float(s1[e + 1]) : -INFINITY);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        mx0 = fmaxf(mx0, __uint_as_float(s0[e]));
                        mx1 = fmaxf(mx1, __uint_as_float(s0[e + 1]));
                        mx2 = fmaxf(mx2, __uint_as_float(s1[e]));
                        mx3 = fmaxf(mx3, __uint_as_float(s1[e + 1]));
                    }
                }
                float m_part = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                if (need_mask && row_empty) m_part = -INFINITY;
                // exchange the half-row maxima (buffer by tile parity).  After this barrier BOTH threads of the row have their
                // scores in registers, so either may overwrite any column of the S row with P.
                xg[(t & 1) * 256 + half * 128 + row] = m_part;
                if (g == 0)
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                else
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                float m_tile = fmaxf(m_part, xg[(t & 1) * 256 + (half ^ 1) * 128 + row]) * sl2;
                const float m_new = fmaxf(m_ref, m_tile);
                float alpha = 1.f;
                if (t > 0) {
                    const bool grow = (m_new - m_ref) > 8.0f;
                    if (__any_sync(0xffffffffu, grow)) {   // both threads of a row decide alike (same m_new, m_ref)
                        if (grow) {
                            alpha = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);
                            m_ref = m_new;
                        }
#pragma unroll 1
                        for (int c = 0; c < OC / 32; ++c) {
                            uint32_t o[32];
                            tmem_ld32(o_col + c * 32, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                            tmem_st32(o_col + c * 32, o);
                        }
                        tmem_st_wait();
                    }
                } else {
                    m_ref = m_new;
                }
                const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                auto exp_pack = [&](auto masked_tag, const uint32_t (&sv)[32], int col0, uint32_t (&pk)[16]) {
                    constexpr bool MASKED = decltype(masked_tag)::value;
#pragma unroll
                    for (int e = 0; e < 32; e += 2) {
                        float p0 = a2_exp2(fmaf(__uint_as_float(sv[e]), sl2, neg_m));
                        float p1 = a2_exp2(fmaf(__uint_as_float(sv[e + 1]), sl2, neg_m));
                        if (MASKED) {
                            p0 = (static_cast<uint32_t>(col0 + e - vlo) <= vspan && !row_empty) ? p0 : 0.f;
                            p1 = (static_cast<uint32_t>(col0 + e + 1 - vlo) <= vspan && !row_empty) ? p1 : 0.f;
                        }
                        if (e & 2) {
                            l2 += p0;
                            l3 += p1;
                        } else {
                            l0 += p0;
                            l1 += p1;
                        }
                        pk[e >> 1] = pack_bf16x2(p0, p1);
                    }
                };
                uint32_t pk[16];
                // P columns [half * 32, half * 32 + 32): keys cbase .. cbase + 63, two per 32-bit cell
                if (need_mask)
                    exp_pack(std::true_type{}, s0, cbase, pk);
                else
                    exp_pack(std::false_type{}, s0, cbase, pk);
                tmem_st16(s_col + half * 32, pk);
                if (need_mask)
                    exp_pack(std::true_type{}, s1, cbase + 32, pk);
                else
                    exp_pack(std::false_type{}, s1, cbase + 32, pk);
                tmem_st16(s_col + half * 32 + 16, pk);
                l_run = l_run * alpha + ((l0 + l1) + (l2 + l3));
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[g]);
            }

            // ---- epilogue: the two threads of a row add their partial sums and split the D output columns
            if (n_g > 0) {
                mbar_wait(&o_full[g], 0);
                tc_fence_after();
            }
            xg[512 + half * 128 + row] = l_run;
            if (g == 0)
                asm volatile("bar.sync 1, 256;" ::: "memory");
            else
                asm volatile("bar.sync 2, 256;" ::: "memory");
            const float l_tot = l_run + xg[512 + (half ^ 1) * 128 + row];
            const bool live = (g == 0) || activeB;
            const float inv_l = (l_tot > 0.f) ? 1.0f / l_tot : 0.f;
            bf16* orow = a.out + (static_cast<size_t>(b) * a.Tq + qi) * a.ldo + h * D + half * OC;
#pragma unroll 1
            for (int c = 0; c < OC / 32; ++c) {
                uint32_t o[32];
                if (n_g > 0) {
                    tmem_ld32(o_col + c * 32, o);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = 0u;
                }
                if (live && qi < a.Tq) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 pk4 = make_uint4(
                            pack_bf16x2(__uint_as_float(o[8 * q4]) * inv_l, __uint_as_float(o[8 * q4 + 1]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * q4 + 2]) * inv_l, __uint_as_float(o[8 * q4 + 3]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * q4 + 4]) * inv_l, __uint_as_float(o[8 * q4 + 5]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * q4 + 6]) * inv_l, __uint_as_float(o[8 * q4 + 7]) * inv_l));
                        reinterpret_cast<uint4*>(orow + c * 32)[q4] = pk4;
                    }
                }
            }
        } else {
        // ---- softmax: group g = Q tile g, one thread per query row (= TMEM lane)
            const int g = (warp - 2) >> 2;
            const int qd = warp & 3;               // TMEM lane quarter this warp may touch
            const int row = qd * 32 + lane;
            const int q0g = q0 + g * A2_BM;
            const int qi = q0g + row;
            const int n_g = g ? nB : nA;
            const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
            const uint32_t s_col = tmem_S0 + g * A2_BN + lane_off;
            const uint32_t o_col = tmem_O0 + g * D + lane_off;
            const int kvl = a.kv_len ? min(a.kv_len[b], a.Tk) : a.Tk;
            const int kvs = a.kv_start ? a.kv_start[b] : 0;
            const int causal_hi = a.causal ? qi + (a.Tk - a.Tq) : 0x7fffffff;  // last visible key (inclusive)
            const float sl2 = a.scale_log2;
            float m_ref = -INFINITY, l_run = 0.f;
    
            for (int t = 0; t < n_g; ++t) {
                const int k0 = (j_lo + t) * A2_BN;
                mbar_wait(&s_full[g], t & 1);
                tc_fence_after();
                const bool need_mask = (k0 < kvs) || (k0 + A2_BN > kvl) || (a.causal && k0 + A2_BN - 1 > q0g + (a.Tk - a.Tq));
                const int vlo = max(kvs - k0, 0);
                const int vhi = min(min(kvl - 1, causal_hi) - k0, A2_BN - 1);
                const uint32_t vspan = static_cast<uint32_t>(vhi - vlo);
                const bool row_empty = vhi < vlo;
                // ---- pass 1: row max (thread-local)
                float m_tile;
                {
                    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
                    // not unrolled over the column chunks: the fully unrolled kernel was 64 KB of SASS and the softmax warps spent 19 %
                    // of their non-waiting samples on instruction-cache misses (stall_no_inst, profiles/r02c_ncu_attention_v2.md)
    #pragma unroll 1
                    for (int c = 0; c < 4; c += 2) {
                        uint32_t v0[32], v1[32];
                        tmem_ld32(s_col + c * 32, v0);
                        tmem_ld32(s_col + c * 32 + 32, v1);
                        tmem_ld_wait();
                        if (need_mask) {
    #pragma unroll
                            for (int e = 0; e < 32; e += 2) {
                                const bool ok0 = static_cast<uint32_t>(c * 32 + e - vlo) <= vspan;
                                const bool ok1 = static_cast<uint32_t>(c * 32 + e + 1 - vlo) <= vspan;
                                const bool ok2 = static_cast<uint32_t>(c * 32 + 32 + e - vlo) <= vspan;
                                const bool ok3 = static_cast<uint32_t>(c * 32 + 33 + e - vlo) <= vspan;
                                mx0 = fmaxf(mx0, ok0 ? __uint_as_float(v0[e]) : -INFINITY);
                                mx1 = fmaxf(mx1, ok1 ? __uint_as_float(v0[e + 1]) : -INFINITY);
                                mx2 = fmaxf(mx2, ok2 ? __uint_as_float(v1[e]) : -INFINITY);
                                mx3 = fmaxf(mx3, ok3 ? __uint_as_float(v1[e + 1]) : -INFINITY);
                            }
                        } else {
    #pragma unroll
                            for (int e = 0; e < 32; e += 2) {
                                mx0 = fmaxf(mx0, __uint_as_float(v0[e]));
                                mx1 = fmaxf(mx1, __uint_as_float(v0[e + 1]));
                                mx2 = fmaxf(mx2, __uint_as_float(v1[e]));
                                mx3 = fmaxf(mx3, __uint_as_float(v1[e + 1]));
                            }
                        }
                    }
                    m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                    if (need_mask && row_empty) m_tile = -INFINITY;
                }
                m_tile *= sl2;  // sl2 > 0: max commutes with the scale
                const float m_new = fmaxf(m_ref, m_tile);
                float alpha = 1.f;
                if (t > 0) {
                    const bool grow = (m_new - m_ref) > 8.0f;  // also true when m_ref == -inf and m_new finite
                    if (__any_sync(0xffffffffu, grow)) {       // warp-uniform: the TMEM accesses below are .sync.aligned
                        // O_g is about to be rewritten: P_g(t-1) V has retired -- S_g(t), whose completion this thread has just
                        // waited for, was issued after it and tcgen05.mma retires in issue order (see the MMA warp)
                        if (grow) {
                            alpha = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);
                            m_ref = m_new;
                        }
    #pragma unroll 1
                        for (int c = 0; c < D / 32; ++c) {
                            uint32_t o[32];
                            tmem_ld32(o_col + c * 32, o);
                            tmem_ld_wait();
    #pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                            tmem_st32(o_col + c * 32, o);
                        }
                        tmem_st_wait();
                    }
                } else {
                    m_ref = m_new;
                }
                const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
                // ---- pass 2: p = 2^(s*scale - m) -> bf16 pairs written back IN PLACE over the already-read part of the S row
                //      (P chunk c lands in columns [16c, 16c+16), all inside S columns [0, 32(c+1)) which this thread has read)
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                auto pass2 = [&](auto masked_tag) {
                    constexpr bool MASKED = decltype(masked_tag)::value;
    #pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld32(s_col + c * 32, v);
                        tmem_ld_wait();
                        uint32_t pk[16];
    #pragma unroll
                        for (int e = 0; e < 32; e += 2) {
                            float p0 = a2_exp2(fmaf(__uint_as_float(v[e]), sl2, neg_m));
                            float p1 = a2_exp2(fmaf(__uint_as_float(v[e + 1]), sl2, neg_m));
                            if (MASKED) {
                                p0 = (static_cast<uint32_t>(c * 32 + e - vlo) <= vspan && !row_empty) ? p0 : 0.f;
                                p1 = (static_cast<uint32_t>(c * 32 + e + 1 - vlo) <= vspan && !row_empty) ? p1 : 0.f;
                            }
                            if (e & 2) {
                                l2 += p0;
                                l3 += p1;
                            } else {
                                l0 += p0;
                                l1 += p1;
                            }
                            pk[e >> 1] = pack_bf16x2(p0, p1);
                        }
                        tmem_st16(s_col + c * 16, pk);
                    }
                };
                if (need_mask)
                    pass2(std::true_type{});
                else
                    pass2(std::false_type{});
                l_run = l_run * alpha + ((l0 + l1) + (l2 + l3));
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[g]);
            }
    
            // ---- epilogue: O / l -> bf16 -> global (one row per thread)
            if (n_g > 0) {
                mbar_wait(&o_full[g], 0);
                tc_fence_after();
            }
            const bool live = (g == 0) || activeB;
            const float inv_l = (l_run > 0.f) ? 1.0f / l_run : 0.f;
            bf16* orow = a.out + (static_cast<size_t>(b) * a.Tq + qi) * a.ldo + h * D;
    #pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
                uint32_t o[32];
                if (n_g > 0) {
                    tmem_ld32(o_col + c * 32, o);
                    tmem_ld_wait();
                } else {
    #pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = 0u;
                }
                if (live && qi < a.Tq) {
    #pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const uint4 pk = make_uint4(
                            pack_bf16x2(__uint_as_float(o[8 * q4]) * inv_l, __uint_as_float(o[8 * q4 + 1]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * q4 + 2]) * inv_l, __uint_as_float(o[8 * q4 + 3]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * q4 + 4]) * inv_l, __uint_as_float(o[8 * q4 + 5]) * inv_l),
                            pack_bf16x2(__uint_as_float(o[8 * q4 + 6]) * inv_l, __uint_as_float(o[8 * q4 + 7]) * inv_l));
                        reinterpret_cast<uint4*>(orow + c * 32)[q4] = pk;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int D, int TPR>
static int launch_attention2(cudaStream_t stream, const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv,
                             const Attn2Args& a, int B) {
    using Cfg = Attn2Cfg<D>;
    auto kern = attention2_kernel<D, TPR>;
    static DeviceOnce once;
    if (once.first()) {
        AF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    }
    dim3 grid(ceil_div(a.Tq, 2 * A2_BM), a.H, B);
    kern<<<grid, 64 + 256 * TPR, Cfg::SMEM_BYTES, stream>>>(mq, mk, mv, a);
    AF3_CHECK_LAUNCH();
    return 0;
}

// Same contract as attention() in attention_tcgen05.cu (which dispatches here).
int attention_v2(cudaStream_t stream, const bf16* q, int ldq, const bf16* k, const bf16* v, int ldk, int kv_layout,
                 int Tk_pitch, bf16* out, int ldo, int B, int H, int Hkv, int D, int Tq, int Tk, float scale, int causal,
                 const int* kv_len, const int* kv_start) {
    Attn2Args a{};
    a.Tq = Tq;
    a.Tk = Tk;
    a.H = H;
    a.Hkv = Hkv;
    a.causal = causal;
    a.kv_layout = kv_layout;
    a.scale_log2 = scale * A2_LOG2E;
    a.kv_len = kv_len;
    a.kv_start = kv_start;
    a.out = out;
    a.ldo = ldo;
    CUtensorMap mq, mk, mv;
    if (int e = make_tmap_3d(&mq, q, (uint64_t)H * D, Tq, B, ldq, (uint64_t)Tq * ldq, 64, A2_BM, 1)) return e;
    if (kv_layout) {
        if (int e = make_tmap_3d(&mk, k, D, Tk, (uint64_t)B * Hkv, ldk, (uint64_t)Tk_pitch * ldk, 64, A2_BN, 1)) return e;
        if (int e = make_tmap_3d(&mv, v, D, Tk, (uint64_t)B * Hkv, ldk, (uint64_t)Tk_pitch * ldk, 64, A2_BN, 1)) return e;
    } else {
        if (int e = make_tmap_3d(&mk, k, (uint64_t)Hkv * D, Tk, B, ldk, (uint64_t)Tk * ldk, 64, A2_BN, 1)) return e;
        if (int e = make_tmap_3d(&mv, v, (uint64_t)Hkv * D, Tk, B, ldk, (uint64_t)Tk * ldk, 64, A2_BN, 1)) return e;
    }
    // softmax threads per query row (see the kernel).  One per row is the default: the two-per-row variant (AF3_ATTN_TPR=2) is
    // parity-clean but measured 7-12 % slower (profiles/r02i_microbench_attention.json: 96 registers with spills at 18 warps, one
    // more named barrier per tile)
    const char* e = getenv("AF3_ATTN_TPR");
    const bool one = !(e && e[0] == '2');
    if (D == 64) return one ? launch_attention2<64, 1>(stream, mq, mk, mv, a, B) : launch_attention2<64, 2>(stream, mq, mk, mv, a, B);
    return one ? launch_attention2<128, 1>(stream, mq, mk, mv, a, B) : launch_attention2<128, 2>(stream, mq, mk, mv, a, B);
}

}  // namespace af3
