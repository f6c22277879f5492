// Log-mel frontend + masked instance norm for sm_100a.
//
// Replaces (reference file:line): normalize_signal models.py:684-686; LogFilterBankFrontend
// .forward models.py:565-597 (pre-emphasis :572, mask :575, reflect/zero pad :577-582, STFT
// :590, power :592-594, mel+eps+log :595); MaskedInstanceNorm1d.forward models.py:694-719 as
// called from JasperNet.forward :298-301.
//
// These stages are HBM/latency bound (57.6 KB per audio second), so there is no
// GEMM here: each warp runs one radix-2 DIF FFT of nfft complex points holding TWO real frames
// (frame A in the real lane, frame B in the imaginary lane), split afterwards with the
// conjugate-symmetry identity.  The mel projection walks only the non-zero band of each
// triangular filter.
#include "common.cuh"
#include "../../include/convasr_b200.h"
#include <atomic>

namespace cab {
extern std::atomic<int64_t> g_launch_count;

// ------------------------------------------------------------------------------------------
// K1: per-utterance max |x|  (models.py:685)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void absmax_kernel(const T* __restrict__ x, int Tn, float* __restrict__ out) {
    const int b = blockIdx.y;
    const T* row = x + size_t(b) * Tn;
    const int base = blockIdx.x * (256 * 16);
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int i = base + k * 256 + threadIdx.x;
        if (i < Tn) m = fmaxf(m, fabsf((float)row[i]));
    }
    m = warp_max(m);
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 8) {
        m = sm[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
        // non-negative floats order like their bit patterns
        if (threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned int*>(out + b), __float_as_uint(m));
    }
}

// ------------------------------------------------------------------------------------------
// K2-K4: framing + window + FFT + power + mel + log
// ------------------------------------------------------------------------------------------
constexpr int kFrontendWarps = 8;
constexpr int kFramesPerCta = 32;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

template <int LOG2N>
__device__ __forceinline__ int bitrev(int v) {
    return (int)(__brev((unsigned)v) >> (32 - LOG2N));
}

// In-warp DIF FFT.  Element n = j*32 + lane lives in z[j] of lane `lane`.
// Result: position p holds X[bitrev(p)].
template <int LOG2N>
__device__ __forceinline__ void warp_fft(float2 (&z)[(1 << LOG2N) / 32], const float2* __restrict__ tw,
                                         int lane) {
    constexpr int N = 1 << LOG2N;
    constexpr int PER = N / 32;
    // in-lane stages: spans N/2 ... 32  (j distance h/32)
#pragma unroll
    for (int h = N / 2; h >= 32; h >>= 1) {
        const int jd = h / 32;
        const int tw_mul = (N / 2) / h;  // W_{2h}^r = W_N^{r * N/(2h)}
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if ((j & jd) == 0) {
                const int r = ((j * 32) & (h - 1)) + lane;  // n mod h, with n = j*32+lane, h >= 32
                float2 a = z[j], b = z[j + jd];
                z[j] = make_float2(a.x + b.x, a.y + b.y);
                float2 d = make_float2(a.x - b.x, a.y - b.y);
                z[j + jd] = cmul(d, tw[r * tw_mul]);
            }
        }
    }
    // cross-lane stages: spans 16 ... 1
#pragma unroll
    for (int h = 16; h >= 1; h >>= 1) {
        const int tw_mul = (N / 2) / h;
        const bool upper = (lane & h) != 0;
        const float2 w = tw[(lane & (h - 1)) * tw_mul];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            float2 mine = z[j];
            float2 other;
            other.x = __shfl_xor_sync(0xffffffffu, mine.x, h);
            other.y = __shfl_xor_sync(0xffffffffu, mine.y, h);
            if (upper) {
                float2 d = make_float2(other.x - mine.x, other.y - mine.y);
                z[j] = cmul(d, w);
            } else {
                z[j] = make_float2(mine.x + other.x, mine.y + other.y);
            }
        }
    }
}

struct FrontendParams {
    const void* signal;
    const float* xlen;
    const float* absmax;
    const float* window;
    const float* mel_fb;
    const int* mel_band;  // [n_mels, 2]: first / one-past-last non-zero bin
    const float2* twiddle;
    float* out;
    int is_int16, B, T, F, win, hop, nfft, n_mels, n_freq;
    int first_off;  // sample index of frame 0, window tap 0:  -(nfft/2) + (nfft-win)/2
    float preemph, log_eps, denom_mult;
    int normalize;
};

template <int LOG2N>
__global__ void __launch_bounds__(kFrontendWarps * 32)
logmel_kernel(const FrontendParams p) {
    constexpr int N = 1 << LOG2N;
    constexpr int PER = N / 32;
    extern __shared__ float smem_f[];
    const int n_samples = (kFramesPerCta - 1) * p.hop + p.win;
    float* s_e = smem_f;                                  // [n_samples] pre-emphasised signal
    float* s_win = s_e + ((n_samples + 3) & ~3);          // [win]
    float2* s_tw = reinterpret_cast<float2*>(s_win + ((p.win + 3) & ~3));  // [N/2]
    float2* s_z = s_tw + N / 2;                           // [warps][N]
    float* s_pw = reinterpret_cast<float*>(s_z + kFrontendWarps * N);  // [warps][2][n_freq_pad]
    const int nfp = (p.n_freq + 3) & ~3;
    float* s_out = s_pw + kFrontendWarps * 2 * nfp;       // [n_mels][kFramesPerCta+1]

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < p.win; i += blockDim.x) s_win[i] = p.window[i];
    for (int i = tid; i < N / 2; i += blockDim.x) s_tw[i] = p.twiddle[i];

    // ---- stage the pre-emphasised, masked, reflect-padded signal
    int n_valid = p.T;
    if (p.xlen != nullptr) n_valid = min(p.T, frac_len(p.xlen[b], p.T));
    float denom = 1.f;
    if (p.normalize) denom = __fmul_rn(__fadd_rn(p.absmax[b], 1e-5f), p.denom_mult);
    const bool reflect = (p.nfft / 2) < p.T;  // models.py:578: 'constant' pad if pad >= T
    const int start = f0 * p.hop + p.first_off;
    const float* xf = static_cast<const float*>(p.signal) + size_t(b) * p.T;
    const short* xs = static_cast<const short*>(p.signal) + size_t(b) * p.T;
    for (int i = tid; i < n_samples; i += blockDim.x) {
        int n = start + i;
        float e = 0.f;
        bool ok = true;
        if (n < 0) { if (reflect) n = -n; else ok = false; }
        if (ok && n < n_valid) {
            float x1 = p.is_int16 ? (float)xs[n] : xf[n];
            if (p.normalize) x1 = __fdiv_rn(x1, denom);
            if (n > 0 && p.preemph > 0.f) {
                float x0 = p.is_int16 ? (float)xs[n - 1] : xf[n - 1];
                if (p.normalize) x0 = __fdiv_rn(x0, denom);
                e = __fsub_rn(x1, __fmul_rn(p.preemph, x0));
            } else {
                e = x1;
            }
        }
        s_e[i] = e;
    }
    __syncthreads();

    float2* zbuf = s_z + warp * N;
    float* pwA = s_pw + warp * 2 * nfp;
    float* pwB = pwA + nfp;

    for (int pair = warp; pair < kFramesPerCta / 2; pair += kFrontendWarps) {
        const int fa = 2 * pair, fb = 2 * pair + 1;  // local frame indices
        float2 z[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int n = j * 32 + lane;
            float2 v = make_float2(0.f, 0.f);
            if (n < p.win) {
                const float w = s_win[n];
                v.x = w * s_e[fa * p.hop + n];
                v.y = w * s_e[fb * p.hop + n];
            }
            z[j] = v;
        }
        warp_fft<LOG2N>(z, s_tw, lane);
#pragma unroll
        for (int j = 0; j < PER; ++j) zbuf[bitrev<LOG2N>(j * 32 + lane)] = z[j];
        __syncwarp();
        // split the two real spectra: A[k] = (Z[k] + conj Z[N-k]) / 2, B[k] = (Z[k] - conj Z[N-k]) / 2i
        for (int k = lane; k < p.n_freq; k += 32) {
            const float2 zk = zbuf[k];
            const float2 zn = zbuf[(N - k) & (N - 1)];
            const float ar = 0.5f * (zk.x + zn.x), ai = 0.5f * (zk.y - zn.y);
            const float br = 0.5f * (zk.y + zn.y), bi = 0.5f * (zn.x - zk.x);
            pwA[k] = ar * ar + ai * ai;
            pwB[k] = br * br + bi * bi;
        }
        __syncwarp();
        for (int m = lane; m < p.n_mels; m += 32) {
            const int lo = p.mel_band[2 * m], hi = p.mel_band[2 * m + 1];
            const float* mrow = p.mel_fb + size_t(m) * p.n_freq;
            float accA = 0.f, accB = 0.f;
            for (int k = lo; k < hi; ++k) {
                const float w = __ldg(mrow + k);
                accA = fmaf(w, pwA[k], accA);
                accB = fmaf(w, pwB[k], accB);
            }
            s_out[m * (kFramesPerCta + 1) + fa] = logf(accA + p.log_eps);
            s_out[m * (kFramesPerCta + 1) + fb] = logf(accB + p.log_eps);
        }
        __syncwarp();
    }
    __syncthreads();
    // coalesced store: each warp writes rows of kFramesPerCta consecutive frames
    for (int m = warp; m < p.n_mels; m += kFrontendWarps) {
        const int f = f0 + lane;
        if (f < p.F) p.out[(size_t(b) * p.n_mels + m) * p.F + f] = s_out[m * (kFramesPerCta + 1) + lane];
    }
}

// ------------------------------------------------------------------------------------------
// K5: masked instance norm statistics + normalise/transpose/pack
// ------------------------------------------------------------------------------------------
// one warp per (b, c): two passes over the (L2-resident) row, exactly the reference's
// mean -> centred sum of squares order (models.py:715-718 / :704-709).
__global__ void instnorm_stats_kernel(const float* __restrict__ feat, const float* __restrict__ xlen,
                                      int B, int C, int F, float eps, float* __restrict__ stats) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= B * C) return;
    const int b = gw / C;
    int n = F;
    if (xlen != nullptr) n = min(F, frac_len(xlen[b], F));
    const float* row = feat + size_t(gw) * F;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s += row[i];
    s = warp_sum(s);
    const float mean = s / (float)n;
    float q = 0.f;
    for (int i = lane; i < n; i += 32) {
        const float d = row[i] - mean;
        q = fmaf(d, d, q);
    }
    q = warp_sum(q);
    const float sd = sqrtf(q / (float)n + eps);
    if (lane == 0) {
        stats[2 * gw] = mean;
        stats[2 * gw + 1] = sd;
    }
}

constexpr int kPackFrames = 32;
__global__ void __launch_bounds__(256)
instnorm_pack_kernel(const float* __restrict__ feat, const float* __restrict__ xlen,
                     const float* __restrict__ stats, int B, int C, int F, int F_pad, int C_pad,
                     __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                     float* __restrict__ out_f32, int normalize) {
    extern __shared__ float tile[];  // [kPackFrames][C_pad + 1]
    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kPackFrames;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n = F;
    if (xlen != nullptr) n = min(F, frac_len(xlen[b], F));
    const int ldt = C_pad + 1;
    for (int c = warp; c < C_pad; c += 8) {
        const int f = f0 + lane;
        float v = 0.f;
        if (c < C && f < n) {
            v = feat[(size_t(b) * C + c) * F + f];
            if (normalize) {
                const float mean = stats[2 * (b * C + c)], sd = stats[2 * (b * C + c) + 1];
                v = __fdiv_rn(v - mean, sd);
            }
        }
        if (out_f32 != nullptr && c < C && f < F) out_f32[(size_t(b) * C + c) * F + f] = v;
        tile[lane * ldt + c] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPackFrames * C_pad; i += blockDim.x) {
        const int fl = i / C_pad, c = i - fl * C_pad;
        const int f = f0 + fl;
        if (f < F_pad) {
            const float v = tile[fl * ldt + c];
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const size_t o = (size_t(b) * F_pad + f) * C_pad + c;
            if (out_hi != nullptr) out_hi[o] = h;
            if (out_lo != nullptr) out_lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
}

}  // namespace cab

using namespace cab;

extern "C" int cab_frontend_logmel(const void* signal, int signal_is_int16, const float* xlen_frac,
                                   int B, int T, int win_length, int hop, int nfft, int n_mels,
                                   const float* window, const float* mel_fb, const int32_t* mel_band,
                                   const float* twiddle, float preemphasis, float log_eps,
                                   int normalize_signal, float denom_multiplier, float* out_logmel,
                                   float* ws_absmax, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(signal && window && mel_fb && mel_band && twiddle && out_logmel, "null pointer argument");
    CAB_CHECK_ARG(B > 0 && T > 0, "bad shape B=%d T=%d", B, T);
    CAB_CHECK_ARG(nfft == 256 || nfft == 512 || nfft == 1024, "nfft=%d unsupported (256/512/1024)", nfft);
    CAB_CHECK_ARG(win_length > 0 && win_length <= nfft && hop > 0, "bad window %d / hop %d", win_length, hop);
    CAB_CHECK_ARG(!normalize_signal || ws_absmax, "ws_absmax required when normalize_signal");
    const int F = T / hop + 1;

    if (normalize_signal) {
        CAB_CHECK_CUDA(cudaMemsetAsync(ws_absmax, 0, sizeof(float) * B, stream));
        dim3 grid((T + 256 * 16 - 1) / (256 * 16), B);
        if (signal_is_int16)
            absmax_kernel<short><<<grid, 256, 0, stream>>>(static_cast<const short*>(signal), T, ws_absmax);
        else
            absmax_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(signal), T, ws_absmax);
        CAB_CHECK_LAUNCH();
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }

    FrontendParams p;
    p.signal = signal; p.xlen = xlen_frac; p.absmax = ws_absmax; p.window = window;
    p.mel_fb = mel_fb; p.mel_band = mel_band; p.twiddle = reinterpret_cast<const float2*>(twiddle);
    p.out = out_logmel; p.is_int16 = signal_is_int16; p.B = B; p.T = T; p.F = F;
    p.win = win_length; p.hop = hop; p.nfft = nfft; p.n_mels = n_mels; p.n_freq = nfft / 2 + 1;
    p.first_off = -(nfft / 2) + (nfft - win_length) / 2;
    p.preemph = preemphasis; p.log_eps = log_eps; p.denom_mult = denom_multiplier;
    p.normalize = normalize_signal;

    const int n_samples = (kFramesPerCta - 1) * hop + win_length;
    const int nfp = (p.n_freq + 3) & ~3;
    size_t smem = sizeof(float) * (((n_samples + 3) & ~3) + ((win_length + 3) & ~3)) +
                  sizeof(float2) * (nfft / 2 + kFrontendWarps * nfft) +
                  sizeof(float) * (kFrontendWarps * 2 * nfp + n_mels * (kFramesPerCta + 1));
    dim3 grid((F + kFramesPerCta - 1) / kFramesPerCta, B);
    auto launch = [&](auto kern) -> int {
        static size_t smem_set = 0;  // one static per kernel instantiation (generic lambda)
        if (smem > smem_set) {
            CAB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            smem_set = smem;
        }
        kern<<<grid, kFrontendWarps * 32, smem, stream>>>(p);
        CAB_CHECK_LAUNCH();
        return 0;
    };
    int rc = nfft == 256 ? launch(logmel_kernel<8>) : nfft == 512 ? launch(logmel_kernel<9>) : launch(logmel_kernel<10>);
    if (rc) return rc;
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

extern "C" int cab_instnorm_pack(const float* feat, const float* xlen_frac, int B, int C, int F,
                                 float eps, int normalize, int F_pad, int C_pad, void* out_hi,
                                 void* out_lo, float* out_f32, float* ws_stats, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(feat && ws_stats, "null pointer argument");
    CAB_CHECK_ARG(out_hi || out_f32, "no output requested");
    CAB_CHECK_ARG(B > 0 && C > 0 && F > 0 && F_pad >= F && C_pad >= C, "bad shape B=%d C=%d F=%d F_pad=%d C_pad=%d", B, C, F, F_pad, C_pad);
    if (normalize) {
        const int warps = B * C;
        const int blocks = (warps * 32 + 255) / 256;
        instnorm_stats_kernel<<<blocks, 256, 0, stream>>>(feat, xlen_frac, B, C, F, eps, ws_stats);
        CAB_CHECK_LAUNCH();
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }
    {
        dim3 grid((F_pad + kPackFrames - 1) / kPackFrames, B);
        size_t smem = sizeof(float) * kPackFrames * (C_pad + 1);
        CAB_CHECK_ARG(smem <= 48 * 1024, "C_pad=%d too large for the pack tile", C_pad);
        instnorm_pack_kernel<<<grid, 256, smem, stream>>>(feat, xlen_frac, ws_stats, B, C, F, F_pad, C_pad,
                                                          static_cast<__nv_bfloat16*>(out_hi),
                                                          static_cast<__nv_bfloat16*>(out_lo), out_f32, normalize);
        CAB_CHECK_LAUNCH();
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
    }
    return 0;
}
