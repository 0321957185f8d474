// Log-mel front end on CUDA cores (fp32), replacing the torch.stft -> |.|^2 -> mel GEMM -> log10 -> floor -> affine
// chain of WhisperFeatureExtractor._torch_extract_fbank_features ([O] WFE:135-164).
//
// One CTA = 64 consecutive frames of one 30 s window.  For each frame (n_fft 400, hop 160, center=True reflect
// padding as torch.stft) the windowed frame y[n] = hann[n]*x[n] is folded with the two symmetries of a real
// 400-point DFT so that only 101 frequency bins x ~100 terms x 4 partial sums are accumulated:
//     a[n] = y[n] + y[400-n], b[n] = y[n] - y[400-n]              (n-fold: cos even, sin odd about n = 200)
//     Ce/Co[k] = sum over even/odd n of a[n] cos(2 pi k n/400), Se/So[k] likewise with b[n] sin(...)
//     Re X[k] = Ce+Co, Re X[200-k] = Ce-Co, |Im X[k]| = |Se+So|, |Im X[200-k]| = |Se-So|      (k-fold, k = 0..100)
// i.e. 40 K MACs per frame instead of 160 K for the plain DFT-as-matmul, and shorter sums (better fp32 error).
// The (cos, sin) table is streamed through shared memory with cp.async; power spectra stay in shared memory;
// the 128 slaney mel filters are applied using their non-zero k ranges; log10(max(.,1e-10)) is written with
// frame-contiguous (coalesced) stores and the per-window maximum is reduced with an atomic.  A second tiny
// kernel applies max(x, winmax-8) and (x+4)/4 in place (the data is still L2 resident).
// Algorithmic HBM bytes: 480000*4 read + 128*3000*4 written = 3.456 MB per window.
#include "common.h"
#include "ptx.cuh"

namespace af3 {

constexpr int NFFT = 400;
constexpr int HOP = 160;
constexpr int NBIN = 201;
constexpr int NMEL = 128;
constexpr int FB = 64;        // frames per CTA
constexpr int FPT = 8;        // frames per warp (8 warps)
constexpr int NH = 201;       // folded n range 0..200
constexpr int KP = 128;       // padded k range (k = 0..100 used)
constexpr int NCHUNK = 8;     // table rows per cp.async stage
constexpr int LM_THREADS = 256;

struct LogmelSmem {
    float2 ab[FB][NH + 1];          // (a[n], b[n]) per frame; later reused as power[f][201]
    float2 tab[2][NCHUNK][KP];      // (cos, sin)(2 pi k n / 400) stage ring
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float ordered_to_float(int v) { return __int_as_float(v >= 0 ? v : v ^ 0x7FFFFFFF); }
__device__ __forceinline__ int float_to_ordered(float f) {
    int v = __float_as_int(f);
    return v >= 0 ? v : v ^ 0x7FFFFFFF;
}

__global__ void __launch_bounds__(LM_THREADS, 1)
logmel_kernel(const float* __restrict__ wave, int n_samples, int n_frames, const float* __restrict__ hann,
              const float2* __restrict__ table /*[201][128]*/, const float* __restrict__ filt /*[201][128]*/,
              const int* __restrict__ klo, const int* __restrict__ khi, float* __restrict__ out, int* __restrict__ win_max) {
    extern __shared__ __align__(16) uint8_t lm_smem[];
    LogmelSmem& s = *reinterpret_cast<LogmelSmem*>(lm_smem);
    const int w = blockIdx.y;
    const int f0 = blockIdx.x * FB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* x = wave + static_cast<size_t>(w) * n_samples;

    auto load_tab = [&](int chunk, int buf) {
        // NCHUNK rows x 128 float2 = 8 KB = 512 x 16 B
        const int n0 = chunk * NCHUNK;
        for (int i = tid; i < NCHUNK * KP / 2; i += LM_THREADS) {
            const int row = i / (KP / 2), c2 = i % (KP / 2);
            const int n = min(n0 + row, NH - 1);
            cp_async16(&s.tab[buf][row][c2 * 2], table + static_cast<size_t>(n) * KP + c2 * 2);
        }
        cp_async_commit();
    };
    load_tab(0, 0);

    // ---- fold the windowed frames:  sample index with torch.stft(center=True, pad_mode="reflect") semantics
    auto sample = [&](int f, int n) -> float {
        int i = f * HOP + n - NFFT / 2;
        if (i < 0) i = -i;
        if (i >= n_samples) i = 2 * (n_samples - 1) - i;
        return __ldg(x + i);
    };
    for (int idx = tid; idx < FB * NH; idx += LM_THREADS) {
        const int fl = idx / NH, n = idx - fl * NH;
        const int f = f0 + fl;
        float a = 0.f, b = 0.f;
        if (f < n_frames) {
            const float y0 = __ldg(hann + n % NFFT) * sample(f, n);  // n = 0..200
            if (n == 0 || n == NFFT / 2) {
                a = y0;
            } else {
                const float y1 = __ldg(hann + NFFT - n) * sample(f, NFFT - n);
                a = y0 + y1;
                b = y0 - y1;
            }
        }
        s.ab[fl][n] = make_float2(a, b);
    }

    // ---- folded DFT: lane owns k = lane + 32 j (j = 0..3), warp owns frames warp*8 .. +7
    float ce[FPT][4], co[FPT][4], se[FPT][4], so[FPT][4];
#pragma unroll
    for (int i = 0; i < FPT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ce[i][j] = co[i][j] = se[i][j] = so[i][j] = 0.f;

    constexpr int NCH = (NH + NCHUNK - 1) / NCHUNK;  // 26 chunks (last partially used)
    for (int ch = 0; ch < NCH; ++ch) {
        if (ch + 1 < NCH) {
            load_tab(ch + 1, (ch + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int buf = ch & 1;
#pragma unroll
        for (int r = 0; r < NCHUNK; r += 2) {
            const int n = ch * NCHUNK + r;  // even n
            if (n < NH) {
                float2 t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = s.tab[buf][r][lane + 32 * j];
#pragma unroll
                for (int i = 0; i < FPT; ++i) {
                    const float2 v = s.ab[warp * FPT + i][n];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        ce[i][j] = fmaf(v.x, t[j].x, ce[i][j]);
                        se[i][j] = fmaf(v.y, t[j].y, se[i][j]);
                    }
                }
            }
            if (n + 1 < NH) {
                float2 t[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) t[j] = s.tab[buf][r + 1][lane + 32 * j];
#pragma unroll
                for (int i = 0; i < FPT; ++i) {
                    const float2 v = s.ab[warp * FPT + i][n + 1];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        co[i][j] = fmaf(v.x, t[j].x, co[i][j]);
                        so[i][j] = fmaf(v.y, t[j].y, so[i][j]);
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- power spectrum into shared memory (reuse the ab region): pw[f][k], row pitch 201 floats
    float* pw = reinterpret_cast<float*>(&s.ab[0][0]);
#pragma unroll
    for (int i = 0; i < FPT; ++i) {
        const int fl = warp * FPT + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = lane + 32 * j;
            if (k <= 100) {
                const float re1 = ce[i][j] + co[i][j], im1 = se[i][j] + so[i][j];
                pw[fl * NBIN + k] = re1 * re1 + im1 * im1;
                if (k < 100) {
                    const float re2 = ce[i][j] - co[i][j], im2 = se[i][j] - so[i][j];
                    pw[fl * NBIN + (200 - k)] = re2 * re2 + im2 * im2;
                }
            }
        }
    }
    __syncthreads();

    // ---- mel filterbank (non-zero k ranges only), log10, store; frame index fastest for coalescing
    float lmax = -INFINITY;
    for (int idx = tid; idx < FB * NMEL; idx += LM_THREADS) {
        const int fl = idx % FB, m = idx / FB;
        const int f = f0 + fl;
        const int k0 = __ldg(klo + m), k1 = __ldg(khi + m);
        float acc = 0.f;
        for (int k = k0; k <= k1; ++k) acc = fmaf(__ldg(filt + k * NMEL + m), pw[fl * NBIN + k], acc);
        const float lg = log10f(fmaxf(acc, 1e-10f));
        if (f < n_frames) {
            out[(static_cast<size_t>(w) * NMEL + m) * n_frames + f] = lg;
            lmax = fmaxf(lmax, lg);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
    if (lane == 0 && lmax > -INFINITY) atomicMax(win_max + w, float_to_ordered(lmax));
}

__global__ void logmel_init_max(int* win_max, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) win_max[i] = float_to_ordered(-INFINITY);
}

// x = (max(x, winmax - 8) + 4) / 4   ([O] WFE:157-161)
__global__ void logmel_finalize(float* __restrict__ out, const int* __restrict__ win_max, int per_win) {
    const int w = blockIdx.y;
    const float floorv = ordered_to_float(win_max[w]) - 8.0f;
    float* p = out + static_cast<size_t>(w) * per_win;
    const int n4 = per_win / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(p)[i];
        v.x = (fmaxf(v.x, floorv) + 4.0f) / 4.0f;
        v.y = (fmaxf(v.y, floorv) + 4.0f) / 4.0f;
        v.z = (fmaxf(v.z, floorv) + 4.0f) / 4.0f;
        v.w = (fmaxf(v.w, floorv) + 4.0f) / 4.0f;
        reinterpret_cast<float4*>(p)[i] = v;
    }
    if (blockIdx.x == 0)
        for (int i = n4 * 4 + threadIdx.x; i < per_win; i += blockDim.x) p[i] = (fmaxf(p[i], floorv) + 4.0f) / 4.0f;
}

int logmel(cudaStream_t stream, const float* wave, int n_win, int n_samples, const float* hann, const float* table,
           const float* filt, const int* klo, const int* khi, float* out, int* win_max) {
    AF3_REQUIRE(n_win > 0 && n_samples >= NFFT && n_samples % HOP == 0, "logmel: n_samples must be a positive multiple of 160");
    const int n_frames = n_samples / HOP;  // torch.stft yields n_frames+1, the reference drops the last (WFE:150)
    static DeviceOnce once;
    if (once.first()) {
        AF3_CHECK_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)sizeof(LogmelSmem)));
    }
    logmel_init_max<<<ceil_div(n_win, 128), 128, 0, stream>>>(win_max, n_win);
    AF3_CHECK_LAUNCH();
    dim3 grid(ceil_div(n_frames, FB), n_win);
    logmel_kernel<<<grid, LM_THREADS, sizeof(LogmelSmem), stream>>>(wave, n_samples, n_frames, hann,
                                                                   reinterpret_cast<const float2*>(table), filt, klo,
                                                                   khi, out, win_max);
    AF3_CHECK_LAUNCH();
    const int per_win = NMEL * n_frames;
    dim3 g2(min(ceil_div(per_win / 4, 256), 64), n_win);
    logmel_finalize<<<g2, 256, 0, stream>>>(out, win_max, per_win);
    AF3_CHECK_LAUNCH();
    return 0;
}

}  // namespace af3
