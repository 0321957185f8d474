// Few-token GEMM WITHOUT split-K for the short-K projections of the decode step (q/k/v and o: K = 3584), sm_100a:
//   out[tok, feat] = epilogue( sum_k X[tok, k] * W[feat, k] ),   n_tok <= 64,  one CTA per 32 weight rows, full K.
// Replaces, for these two projections, the split-K form of gemm_tcgen05.cu ([O] Q2M:199-202 behind F.linear).
//
// Why: the in-graph timeline (profiles/r02h_decode_timeline_*.md) puts 9 us of every 16 us q/k/v or o projection into the split-K
// TAIL -- partial tile to L2, fence, counter, the last CTA's reload and sum: three dependent L2 round trips while the next kernel's
// weight prefetch saturates HBM -- and only 6 us into the stream.  A 128-row UMMA tile forces the split (28 / 36 row tiles for 148
// SMs); turning the operands around does not: with the ACTIVATIONS as the M operand (tokens = TMEM lanes) the weight tile is the
// N operand, and N = 32 rows per CTA gives 112 / 144 CTAs that each own their outputs outright.  No workspace, no counters, no
// reduction; the accumulator is already token-major, so the epilogue is one tcgen05.ld and four 16-byte stores per token.
//
//   warp 0   weight producer: its whole ring (40 x 4 KB, 160 KB of the CTA's 224 KB slice) is requested BEFORE
//            griddepcontrol.wait -- weights do not depend on earlier kernels -- the rest as stages drain
//   warp 2   activation producer: after the dependency resolves, streams the [n_tok x K] activations (L2 hits: they were
//            just written) through its own 16-stage ring; every CTA reads all of them
//   warp 1   MMA issuer: UMMA 128 x 32 x 16, A = activation tile (rows >= n_tok of the 128-row operand alias whatever follows
//            in shared memory: garbage rows only feed accumulator lanes nobody reads), B = 32 weight rows
//   warp 4+  epilogue, thread = token: bias / residual, or rotary embedding + KV-cache append for the q/k/v projection
//            ([O] Q2M:100-146, CACHE:119-120).  For RoPE a CTA owns 16 rotation PAIRS (d, d + 64) of one head: its 32 weight
//            rows are two 16-row TMA boxes, so both partners of a pair sit in one thread's registers.
#include "common.h"
#include "gemm_epi.h"
#include "ptx.cuh"

#include <cstdlib>

namespace af3 {

struct SkinnyArgs {
    int K, n_tok, n_feat;
    bf16* out;
    int ldo;
    const bf16* bias;
    const bf16* resid;
    int ld_res;
    int flags;
    const float2* rope_cs;  // [n_tok][64] (cos, sin), bf16-rounded
    bf16* k_cache;
    bf16* v_cache;
    const int* rope_pos;
    int rope_H, rope_Hkv, rope_Tmax;
    unsigned long long* trace;
};

template <int NT>
struct SkinnyCfg {
    static constexpr int A_BYTES = NT * 128;  // [NT tokens][64 bf16], 128-byte swizzle
    static constexpr int W_BYTES = 32 * 128;  // [32 weight rows][64 bf16]
    static constexpr int SA = (NT == 32) ? 16 : 12;
    static constexpr int SW = (NT == 32) ? 40 : 32;
    static constexpr int BAR_BYTES = (2 * SA + 2 * SW + 1) * 8 + 16;
    static constexpr int SMEM_BYTES = SA * A_BYTES + SW * W_BYTES + 1024 /*align slack*/ + 1024 /*barriers + TMEM slot*/;
    // the 128-row A operand of the LAST activation stage reads 16 KB from its base: the weight ring behind it must cover that
    static_assert(SW * W_BYTES >= 16384 - A_BYTES, "A-operand over-read must stay inside the allocation");
    static_assert(BAR_BYTES <= 1024, "barrier area");
    static_assert(SMEM_BYTES <= 232448, "shared memory");
};

template <int NT>
__global__ void __launch_bounds__(128 + NT, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const SkinnyArgs a) {
    using Cfg = SkinnyCfg<NT>;
    constexpr int SA = Cfg::SA, SW = Cfg::SW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + SA * Cfg::A_BYTES;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(sW + SW * Cfg::W_BYTES);
    uint64_t* empty_a = full_a + SA;
    uint64_t* full_w = empty_a + SA;
    uint64_t* empty_w = full_w + SW;
    uint64_t* tfull = empty_w + SW;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) trace_mark(a.trace, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_x);
        for (int i = 0; i < SA; ++i) {
            mbar_init(&full_a[i], 1);
            mbar_init(&empty_a[i], 1);
        }
        for (int i = 0; i < SW; ++i) {
            mbar_init(&full_w[i], 1);
            mbar_init(&empty_w[i], 1);
        }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();

    const int nkb = a.K / 64;
    const bool rope = a.flags & EPI_ROPE;
    const int b = blockIdx.x;
    // weight rows of this CTA: two boxes of 16.  RoPE: pairs (d, d + 64) of head b / 4, d = (b % 4) * 16 ...; else 32 consecutive rows
    const int row_lo = rope ? (b >> 2) * 128 + (b & 3) * 16 : b * 32;
    const int row_hi = rope ? row_lo + 64 : row_lo + 16;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % SW, u = kb / SW;
                if (u > 0) mbar_wait(&empty_w[s], (u - 1) & 1);
                mbar_arrive_expect_tx(&full_w[s], Cfg::W_BYTES);
                tma_load_2d(sW + s * Cfg::W_BYTES, &map_w, &full_w[s], kb * 64, row_lo);
                tma_load_2d(sW + s * Cfg::W_BYTES + 2048, &map_w, &full_w[s], kb * 64, row_hi);
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            pdl_wait();
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % SA, u = kb / SA;
                if (u > 0) mbar_wait(&empty_a[s], (u - 1) & 1);
                mbar_arrive_expect_tx(&full_a[s], Cfg::A_BYTES);
                tma_load_2d(sA + s * Cfg::A_BYTES, &map_x, &full_a[s], kb * 64, 0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
            for (int kb = 0; kb < nkb; ++kb) {
                const int sa = kb % SA, sw = kb % SW;
                mbar_wait(&full_w[sw], (kb / SW) & 1);
                mbar_wait(&full_a[sa], (kb / SA) & 1);
                tc_fence_after();
                const uint32_t aA = smem_u32(sA + sa * Cfg::A_BYTES);
                const uint32_t aW = smem_u32(sW + sw * Cfg::W_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_ss(tmem_base, make_smem_desc_sw128(aA + k * 32, 0, 1024), make_smem_desc_sw128(aW + k * 32, 0, 1024), idesc,
                                 (kb > 0) || (k != 0));
                if (kb + SA < nkb) umma_commit(&empty_a[sa]);
                if (kb + SW < nkb) umma_commit(&empty_w[sw]);
            }
            umma_commit(tfull);
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;           // == warp & 3: the TMEM lane quarter this warp may read
        const int tok = ew * 32 + lane;
        const bool tok_ok = tok < a.n_tok;
        const int flags = a.flags;
        pdl_wait();
        if (threadIdx.x == 128) trace_mark(a.trace, 1);
        // everything the epilogue needs besides the accumulator is requested before the accumulator wait
        uint4 rv[4];
        uint4 bv[4];
        float4 csv[8];
        int slot = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            rv[g] = make_uint4(0, 0, 0, 0);
            bv[g] = make_uint4(0, 0, 0, 0);
        }
        if (flags & EPI_BIAS) {
            const uint4* bl = reinterpret_cast<const uint4*>(a.bias + row_lo);
            const uint4* bh = reinterpret_cast<const uint4*>(a.bias + row_hi);
            bv[0] = __ldg(bl);
            bv[1] = __ldg(bl + 1);
            bv[2] = __ldg(bh);
            bv[3] = __ldg(bh + 1);
        }
        if ((flags & EPI_RESID) && tok_ok) {
            const uint4* rp = reinterpret_cast<const uint4*>(a.resid + static_cast<size_t>(tok) * a.ld_res + row_lo);
#pragma unroll
            for (int g = 0; g < 4; ++g) rv[g] = __ldcg(rp + g);
        }
        const int head = b >> 2;
        const bool rotate = rope && head < a.rope_H + a.rope_Hkv;
        if (rotate && tok_ok) {
            const float4* cp = reinterpret_cast<const float4*>(a.rope_cs + static_cast<size_t>(tok) * 64 + (b & 3) * 16);
#pragma unroll
            for (int g = 0; g < 8; ++g) csv[g] = __ldg(cp + g);
        }
        if (rope && head >= a.rope_H) slot = *a.rope_pos;

        mbar_wait(tfull, 0);
        tc_fence_after();
        if (threadIdx.x == 128) trace_mark(a.trace, 2);
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16), v);
        tmem_ld_wait();
        if (tok_ok) {
            const __nv_bfloat16* bh = reinterpret_cast<const __nv_bfloat16*>(bv);
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float x = __uint_as_float(v[j]);
                if (flags & EPI_BIAS) x += __bfloat162float(bh[j]);
                x = bf16_round(x);  // nn.Linear output is bf16
                if (flags & EPI_GELU) x = bf16_round(gelu_erf(x));
                y[j] = x;
            }
            if (!rope) {
                const __nv_bfloat16* rh = reinterpret_cast<const __nv_bfloat16*>(rv);
                uint32_t o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float lo = y[2 * j], hi = y[2 * j + 1];
                    if (flags & EPI_RESID) {
                        lo += __bfloat162float(rh[2 * j]);
                        hi += __bfloat162float(rh[2 * j + 1]);
                    }
                    o[j] = pack_bf16x2(lo, hi);
                }
                uint4* op = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(tok) * a.ldo + row_lo);
#pragma unroll
                for (int g = 0; g < 4; ++g) op[g] = make_uint4(o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
            } else {
                // columns 0..15 = features d = (b % 4) * 16 + j of the head, columns 16..31 = their rotation partners d + 64
                float lo[16], hi[16];
                if (rotate) {
                    const float2* cs = reinterpret_cast<const float2*>(csv);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // rotate_half: out[d] = x[d] cos - x[d + 64] sin,  out[d + 64] = x[d + 64] cos + x[d] sin, each product
                        // rounded to bf16 before the add (the reference multiplies and adds bf16 tensors, [O] Q2M:139-146)
                        lo[j] = bf16_round(y[j] * cs[j].x) + bf16_round(-1.f * y[16 + j] * cs[j].y);
                        hi[j] = bf16_round(y[16 + j] * cs[j].x) + bf16_round(y[j] * cs[j].y);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        lo[j] = y[j];
                        hi[j] = y[16 + j];
                    }
                }
                uint32_t ol[8], oh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    ol[j] = pack_bf16x2(lo[2 * j], lo[2 * j + 1]);
                    oh[j] = pack_bf16x2(hi[2 * j], hi[2 * j + 1]);
                }
                const int d0 = (b & 3) * 16;
                bf16* op;
                bool store = true;
                if (head < a.rope_H) {
                    op = a.out + static_cast<size_t>(tok) * a.ldo + head * 128 + d0;
                } else {
                    const int hk = (head - a.rope_H) % a.rope_Hkv;
                    bf16* cache = (head < a.rope_H + a.rope_Hkv) ? a.k_cache : a.v_cache;
                    // a full cache must never be written past its allocation (the host raises before launching; this guards graph
                    // replays whose position lives on the device)
                    store = slot >= 0 && slot < a.rope_Tmax;
                    op = cache + ((static_cast<size_t>(tok) * a.rope_Hkv + hk) * a.rope_Tmax + (store ? slot : 0)) * 128 + d0;
                }
                if (store) {
                    uint4* p0 = reinterpret_cast<uint4*>(op);
                    uint4* p1 = reinterpret_cast<uint4*>(op + 64);
                    p0[0] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
                    p0[1] = make_uint4(ol[4], ol[5], ol[6], ol[7]);
                    p1[0] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
                    p1[1] = make_uint4(oh[4], oh[5], oh[6], oh[7]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) trace_mark(a.trace, 3);
    if (warp == 1) tmem_dealloc(tmem_base, 32);
}

template <int NT>
static int launch_skinny(const CUtensorMap& mw, const CUtensorMap& mx, const SkinnyArgs& a, cudaStream_t stream) {
    using Cfg = SkinnyCfg<NT>;
    auto kern = gemm_skinny_kernel<NT>;
    static DeviceOnce once;
    if (once.first()) AF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    AF3_CHECK_CUDA(launch_kernel(kern, dim3(a.n_feat / 32), dim3(128 + NT), Cfg::SMEM_BYTES, stream, mw, mx, a));
    return 0;
}

// Largest K taken by the no-split path (AF3_SKINNY_MAXK overrides, 0 turns the path off): every CTA re-reads ALL activations, so
// the long-K down projection (18944) stays with the split-K kernel.
static int skinny_max_k() {
    const char* e = getenv("AF3_SKINNY_MAXK");
    return e ? atoi(e) : 0;   // OFF by default: measured 2x slower (profiles/r02k_microbench_fewtoken_nosplit.json)
}

bool gemm_skinny_applies(int n_tok, int n_feat, int K, int flags, int ldx, int ldw, int ldo, int ld_res, int res_period, const void* x,
                         const void* w, const void* out, const void* bias, const void* resid) {
    if (n_tok > 64 || (n_feat % 32) != 0 || n_feat / 32 > sm_count() || (K % 64) != 0 || K > skinny_max_k()) return false;
    if (flags & ~(EPI_BIAS | EPI_GELU | EPI_RESID | EPI_ROPE)) return false;
    if ((ldo % 8) != 0 || (ldx % 8) != 0 || (ldw % 8) != 0) return false;
    if ((flags & EPI_RESID) && ((ld_res % 8) != 0 || res_period != 0 || (reinterpret_cast<uintptr_t>(resid) & 15))) return false;
    if ((flags & EPI_BIAS) && (reinterpret_cast<uintptr_t>(bias) & 15)) return false;
    if ((reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return false;
    return true;
}

int gemm_skinny(cudaStream_t stream, const bf16* x, int ldx, const bf16* w, int ldw, bf16* out, int ldo, int n_tok, int n_feat, int K,
                int flags, const bf16* bias, const bf16* resid, int ld_res, const RopeEpilogue* rope) {
    SkinnyArgs a{};
    a.K = K;
    a.n_tok = n_tok;
    a.n_feat = n_feat;
    a.out = out;
    a.ldo = ldo;
    a.bias = bias;
    a.resid = resid;
    a.ld_res = ld_res;
    a.flags = flags;
    a.trace = trace_next_slot();
    if (flags & EPI_ROPE) {
        AF3_REQUIRE(rope && !(flags & (EPI_RESID | EPI_GELU)), "gemm: EPI_ROPE is the fused q/k/v projection (bias only)");
        AF3_REQUIRE((reinterpret_cast<uintptr_t>(rope->cs) & 15) == 0 && (reinterpret_cast<uintptr_t>(rope->k_cache) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(rope->v_cache) & 15) == 0,
                    "gemm: rope table and KV cache must be 16-byte aligned");
        a.rope_cs = reinterpret_cast<const float2*>(rope->cs);
        a.k_cache = rope->k_cache;
        a.v_cache = rope->v_cache;
        a.rope_pos = rope->pos;
        a.rope_H = rope->H;
        a.rope_Hkv = rope->Hkv;
        a.rope_Tmax = rope->Tmax;
    }
    CUtensorMap mw, mx;
    if (int e = make_tmap_2d(&mw, w, K, n_feat, ldw, 64, 16)) return e;
    if (n_tok <= 32) {
        if (int e = make_tmap_2d(&mx, x, K, n_tok, ldx, 64, 32)) return e;
        return launch_skinny<32>(mw, mx, a, stream);
    }
    if (int e = make_tmap_2d(&mx, x, K, n_tok, ldx, 64, 64)) return e;
    return launch_skinny<64>(mw, mx, a, stream);
}

}  // namespace af3
