// bf16 GEMM for every nn.Linear / conv-as-GEMM on the AF3 path, written for sm_100a:
//   out[tok, feat] = epilogue( sum_k X[tok, k] * W[feat, k] )        (both operands K-major = nn.Linear layout)
// Replaces the cuBLASLt calls behind  F.linear / F.conv1d  in the reference path
// ([O] AF3M:111-114,141,153-154,204-205,343-344,398-402; Q2M:41-43,46-48,199-202,474-475).
//
// Design (one persistent CTA per SM, 6 warps, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 2-D loads of 128B-swizzled [rows x 64] bf16 tiles into a
//               STAGES-deep smem ring, completion on mbarriers.
//   warp 1      MMA issuer: one thread issues tcgen05.mma (UMMA 128 x BN x 16, bf16 -> fp32) with smem
//               descriptors; accumulators live in TMEM, double-buffered so the epilogue of tile i overlaps the
//               main loop of tile i+1.  tcgen05.commit releases smem stages / publishes accumulators.
//   warps 2-5   epilogue: tcgen05.ld 32 lanes x 32 columns per warp, fused bias / GELU(erf) / residual /
//               SwiGLU with the reference's bf16 rounding points, vectorised global stores.
// "Row operand" = the matrix whose rows become TMEM lanes (128 per UMMA), "col operand" = the one whose rows
// become TMEM columns (BN).  Normal mode: rows = tokens, cols = features.  Swap mode (few tokens, decode):
// rows = weight features (so every UMMA streams 128 weight rows), cols = tokens, and the epilogue writes the
// transposed tile.  Tile order is grouped (GROUP_R row tiles share a col tile) so the 126 MB L2 holds the
// working set of one wave and HBM sees each operand ~once.
#include "common.h"
#include "ptx.cuh"

namespace af3 {

enum : int { EPI_BIAS = 1, EPI_GELU = 2, EPI_RESID = 4, EPI_SWIGLU = 8, EPI_F32OUT = 16 };

struct GemmArgs {
    int R, C, K;  // extents of row operand, col operand, reduction
    int num_r_tiles, num_c_tiles, group_r;
    int n_tok, n_feat;  // logical output extents (features after SwiGLU halving)
    void* out;
    int ldo;
    const bf16* bias;
    const bf16* resid;
    int ld_res;
    int res_period;  // residual row = tok % res_period (positional-embedding add); 0 = plain
    int flags;
};

constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle atom row

template <int BN, int NA, int STAGES>
struct GemmCfg {
    static constexpr int R_BYTES = 128 * NA * BK * 2;
    static constexpr int C_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = R_BYTES + C_BYTES;
    static constexpr int ACC_COLS = NA * BN;
    static constexpr int TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64 : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(2 * ACC_COLS <= 512, "TMEM budget");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
};

__device__ __forceinline__ void tile_coords(int t, const GemmArgs& a, int& r, int& c) {
    const int per_group = a.group_r * a.num_c_tiles;
    const int g = t / per_group;
    const int rem = t - g * per_group;
    const int r0 = g * a.group_r;
    const int gr = min(a.group_r, a.num_r_tiles - r0);
    c = rem / gr;
    r = r0 + (rem - c * gr);
}

// epilogue math for one output element; x is the fp32 accumulator
__device__ __forceinline__ float epi_elem(float x, int flags, float bias, float res) {
    if (flags & EPI_BIAS) x += bias;
    x = bf16_round(x);  // nn.Linear output is bf16
    if (flags & EPI_GELU) x = bf16_round(gelu_erf(x));
    if (flags & EPI_RESID) x = bf16_round(x + res);
    return x;
}
__device__ __forceinline__ float epi_swiglu(float g, float u) {
    g = bf16_round(g);
    u = bf16_round(u);
    const float s = bf16_round(g / (1.0f + __expf(-g)));  // F.silu in bf16: fp32 math, one rounding
    return bf16_round(s * u);
}

template <int BN, int NA, int STAGES, bool SWAP>
__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_c, const GemmArgs a) {
    using Cfg = GemmCfg<BN, NA, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_r);
        tma_prefetch_desc(&map_c);
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 128);
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

    const int num_tiles = a.num_r_tiles * a.num_c_tiles;
    const int k_blocks = (a.K + BK - 1) / BK;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int r, c;
                tile_coords(t, a, r, c);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sR = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sC = sR + Cfg::R_BYTES;
                    mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
#pragma unroll
                    for (int na = 0; na < NA; ++na)
                        tma_load_2d(sR + na * 128 * BK * 2, &map_r, &full[stage], kb * BK, (r * NA + na) * 128);
                    tma_load_2d(sC, &map_c, &full[stage], kb * BK, c * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sR = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t sC = sR + Cfg::R_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t bdesc = make_smem_desc_sw128(sC + k * 32, 0, 1024);
#pragma unroll
                        for (int na = 0; na < NA; ++na) {
                            const uint64_t adesc = make_smem_desc_sw128(sR + na * 128 * BK * 2 + k * 32, 0, 1024);
                            umma_bf16_ss(d_tmem + na * BN, adesc, bdesc, idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[acc]);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int row_in_tile = q * 32 + lane;
        const int flags = a.flags;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int r, c;
            tile_coords(t, a, r, c);
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * Cfg::ACC_COLS;

            if constexpr (!SWAP) {
                const int tok = r * 128 + row_in_tile;
                const bool row_ok = tok < a.n_tok;
                const int res_row = (a.res_period > 0) ? (tok % a.res_period) : tok;
                if (flags & EPI_SWIGLU) {
                    // cols [0, BN/2) = gate, [BN/2, BN) = up of the same BN/2 output features
                    constexpr int HALF = BN / 2;
#pragma unroll 1
                    for (int ch = 0; ch < HALF / 32; ++ch) {
                        uint32_t g[32], u[32];
                        tmem_ld32(taddr + ch * 32, g);
                        tmem_ld32(taddr + HALF + ch * 32, u);
                        tmem_ld_wait();
                        const int f0 = c * HALF + ch * 32;
                        if (row_ok && f0 < a.n_feat) {
                            bf16* o = reinterpret_cast<bf16*>(a.out) + static_cast<size_t>(tok) * a.ldo + f0;
                            if (f0 + 32 <= a.n_feat && (a.ldo & 7) == 0) {
                                uint32_t w[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    w[j] = pack_bf16x2(epi_swiglu(__uint_as_float(g[2 * j]), __uint_as_float(u[2 * j])),
                                                       epi_swiglu(__uint_as_float(g[2 * j + 1]), __uint_as_float(u[2 * j + 1])));
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    reinterpret_cast<uint4*>(o)[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
                            } else {
                                _Pragma("unroll") for (int j = 0; j < 32; ++j) if (f0 + j < a.n_feat)
                                    o[j] = __float2bfloat16_rn(epi_swiglu(__uint_as_float(g[j]), __uint_as_float(u[j])));
                            }
                        }
                    }
                } else {
#pragma unroll 1
                    for (int ch = 0; ch < BN / 32; ++ch) {
                        uint32_t v[32];
                        tmem_ld32(taddr + ch * 32, v);
                        tmem_ld_wait();
                        const int f0 = c * BN + ch * 32;
                        if (row_ok && f0 < a.n_feat) {
                            const bool fullv = (f0 + 32 <= a.n_feat);
                            float x[32];
                            if (fullv && (flags & EPI_BIAS)) {
                                const uint4* bp = reinterpret_cast<const uint4*>(a.bias + f0);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    uint4 bv = __ldg(bp + j);
                                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&bv);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        float2 f = __bfloat1622float2(h[e]);
                                        x[j * 8 + e * 2] = f.x;
                                        x[j * 8 + e * 2 + 1] = f.y;
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    x[j] = ((flags & EPI_BIAS) && f0 + j < a.n_feat) ? __bfloat162float(a.bias[f0 + j]) : 0.f;
                            }
                            float rs[32];
                            if (flags & EPI_RESID) {
                                const bf16* rp = a.resid + static_cast<size_t>(res_row) * a.ld_res + f0;
                                if (fullv && (a.ld_res & 7) == 0) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        uint4 rv = *(reinterpret_cast<const uint4*>(rp) + j);
                                        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&rv);
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            float2 f = __bfloat1622float2(h[e]);
                                            rs[j * 8 + e * 2] = f.x;
                                            rs[j * 8 + e * 2 + 1] = f.y;
                                        }
                                    }
                                } else {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) rs[j] = (f0 + j < a.n_feat) ? __bfloat162float(rp[j]) : 0.f;
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) rs[j] = 0.f;
                            }
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] = epi_elem(__uint_as_float(v[j]), flags, x[j], rs[j]);

                            if (flags & EPI_F32OUT) {
                                float* o = reinterpret_cast<float*>(a.out) + static_cast<size_t>(tok) * a.ldo + f0;
                                if (fullv && (a.ldo & 3) == 0) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j)
                                        reinterpret_cast<float4*>(o)[j] = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                                } else {
                                    _Pragma("unroll") for (int j = 0; j < 32; ++j) if (f0 + j < a.n_feat) o[j] = x[j];
                                }
                            } else {
                                bf16* o = reinterpret_cast<bf16*>(a.out) + static_cast<size_t>(tok) * a.ldo + f0;
                                if (fullv && (a.ldo & 7) == 0) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        reinterpret_cast<uint4*>(o)[j] =
                                            make_uint4(pack_bf16x2(x[8 * j], x[8 * j + 1]), pack_bf16x2(x[8 * j + 2], x[8 * j + 3]),
                                                       pack_bf16x2(x[8 * j + 4], x[8 * j + 5]), pack_bf16x2(x[8 * j + 6], x[8 * j + 7]));
                                } else {
                                    _Pragma("unroll") for (int j = 0; j < 32; ++j) if (f0 + j < a.n_feat) o[j] = __float2bfloat16_rn(x[j]);
                                }
                            }
                        }
                    }
                }
            } else {
                // swapped: TMEM lane = output feature, TMEM column = token
                const int feat = r * 128 + row_in_tile;
                const bool feat_ok = feat < a.n_feat;
                const float bias = ((flags & EPI_BIAS) && feat_ok) ? __bfloat162float(a.bias[feat]) : 0.f;
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ++ch) {
                    uint32_t v[32], u[32];
                    tmem_ld32(taddr + ch * 32, v);
                    if constexpr (NA == 2) tmem_ld32(taddr + BN + ch * 32, u);
                    tmem_ld_wait();
                    const int tok0 = c * BN + ch * 32;
                    if (feat_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int tok = tok0 + j;
                            if (tok < a.n_tok) {
                                float x;
                                if constexpr (NA == 2) {
                                    x = epi_swiglu(__uint_as_float(v[j]), __uint_as_float(u[j]));
                                } else {
                                    float res = 0.f;
                                    if (flags & EPI_RESID) {
                                        const int rr = (a.res_period > 0) ? (tok % a.res_period) : tok;
                                        res = __bfloat162float(a.resid[static_cast<size_t>(rr) * a.ld_res + feat]);
                                    }
                                    x = epi_elem(__uint_as_float(v[j]), flags, bias, res);
                                }
                                if (flags & EPI_F32OUT)
                                    reinterpret_cast<float*>(a.out)[static_cast<size_t>(tok) * a.ldo + feat] = x;
                                else
                                    reinterpret_cast<bf16*>(a.out)[static_cast<size_t>(tok) * a.ldo + feat] = __float2bfloat16_rn(x);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tempty[acc]);
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

template <int BN, int NA, int STAGES, bool SWAP>
static int launch(const CUtensorMap& mr, const CUtensorMap& mc, const GemmArgs& a, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, NA, STAGES>;
    auto kern = gemm_kernel<BN, NA, STAGES, SWAP>;
    static bool configured = false;
    if (!configured) {
        AF3_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        configured = true;
    }
    const int tiles = a.num_r_tiles * a.num_c_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, 192, Cfg::SMEM_BYTES, stream>>>(mr, mc, a);
    AF3_CHECK_LAUNCH();
    return 0;
}

// Host entry.  x: [n_tok, K] bf16 (pitch ldx), w: [n_rows_w, K] bf16 (pitch ldw) where n_rows_w = n_feat, or the
// gate/up-interleaved 2*ceil(n_feat/128)*128 rows when EPI_SWIGLU.  out: [n_tok, n_feat] (pitch ldo).
int gemm_bf16(cudaStream_t stream, const bf16* x, int ldx, const bf16* w, int ldw, void* out, int ldo, int n_tok,
              int n_feat, int K, int flags, const bf16* bias, const bf16* resid, int ld_res, int res_period) {
    AF3_REQUIRE(n_tok > 0 && n_feat > 0 && K > 0, "gemm: empty problem");
    AF3_REQUIRE((K % 8) == 0 && (ldx % 8) == 0 && (ldw % 8) == 0, "gemm: K and pitches must be multiples of 8");
    AF3_REQUIRE(!(flags & EPI_BIAS) || bias, "gemm: bias flag without pointer");
    AF3_REQUIRE(!(flags & EPI_RESID) || resid, "gemm: residual flag without pointer");
    const bool swiglu = flags & EPI_SWIGLU;
    const int w_rows = swiglu ? 2 * ceil_div(n_feat, 128) * 128 : n_feat;
    GemmArgs a{};
    a.K = K;
    a.n_tok = n_tok;
    a.n_feat = n_feat;
    a.out = out;
    a.ldo = ldo;
    a.bias = bias;
    a.resid = resid;
    a.ld_res = ld_res;
    a.res_period = res_period;
    a.flags = flags;
    CUtensorMap mx, mw;
    const bool swap = n_tok <= 64;
    if (!swap) {
        constexpr int BN = 256;
        a.R = n_tok;
        a.C = w_rows;
        a.num_r_tiles = ceil_div(n_tok, 128);
        a.num_c_tiles = ceil_div(w_rows, BN);
        a.group_r = 16;
        if (int e = make_tmap_2d(&mx, x, K, n_tok, ldx, BK, 128)) return e;
        if (int e = make_tmap_2d(&mw, w, K, w_rows, ldw, BK, BN)) return e;
        return launch<BN, 1, 4, false>(mx, mw, a, stream);
    }
    constexpr int BN = 32;
    a.R = w_rows;
    a.C = n_tok;
    a.num_c_tiles = ceil_div(n_tok, BN);
    a.group_r = 1;
    if (int e = make_tmap_2d(&mw, w, K, w_rows, ldw, BK, 128)) return e;
    if (int e = make_tmap_2d(&mx, x, K, n_tok, ldx, BK, BN)) return e;
    if (swiglu) {
        a.num_r_tiles = ceil_div(w_rows, 256);
        return launch<BN, 2, 6, true>(mw, mx, a, stream);
    }
    a.num_r_tiles = ceil_div(w_rows, 128);
    return launch<BN, 1, 8, true>(mw, mx, a, stream);
}

}  // namespace af3
