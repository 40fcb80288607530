// GTCRN-Micro inference kernels for sm_100a (B200).  Hand-written CUDA, fp32 throughout.
//
// Data layout in HBM
//   wav        [B][L]                 fp32
//   spec       [B][T][257][2]         fp32 (internal; user tensors are addressed with element strides)
//   activation [B][T][F][16]          fp32, channel-last: one position's 16 channels = 64 contiguous bytes
//                                     (F = 65 after en0, 33 everywhere else)
//   weights    one packed blob (blob_layout.h): BN folded, every contraction stored [k][ci][co]
//
// Kernel design: the layers are tiny channel contractions (K = 8..144, N = 16) applied to
// B*T*F independent positions, so the path is FP32-FMA bound, not HBM or tensor-core bound (a tcgen05 3xTF32
// prototype was measured and rejected: tools/microbench/umma_tf32_test.cu).  Every contraction uses the same
// register tile: one thread owns P frame slots x 16 (or 8) output channels at one frequency bin; activations come
// from registers or from shared memory as LDS.128 in a [row][ci/4][bin][4] layout (consecutive bins -> consecutive
// 16 bytes, parity-split for the stride-2 convs), and the warp-uniform weights from the constant bank (kernel
// parameters: LDCU + FFMA R,R,UR,R) or, in the GTCN kernel, from shared memory as broadcast LDS.128.
//
// Temporal layers (GTConv 3x3 + TRALite, dilated TCNs) are computed by CTAs / warps that WALK along time over one
// utterance segment, carrying their causal history (2 frames of point-conv output, 2 TRA energies, 2d frames per
// TCN) in shared-memory rings between tiles -- the same role the reference's streaming caches play
// (gtcrn_micro_stream.py:416-429, :490-500) -- so no halo is recomputed except a short warm-up when an utterance
// is split into several segments for occupancy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <utility>

#include <cuda_fp16.h>

#include "blob_layout.h"
#include "blob_full_layout.h"
#include "tc_util.cuh"

namespace gk {

constexpr int kBins = 257;   // n_fft/2 + 1
constexpr int kErb = 129;    // 65 low bins + 64 ERB bands
constexpr int kF1 = 65;      // after en0
constexpr int kF2 = 33;      // after en1
constexpr int kSpecFrame = kBins * 2;
constexpr int kHop = 256;
constexpr int kNfft = 512;

// PReLU(v) = v > 0 ? v : a v.  For a <= 1 this is max(v, a v), for a > 1 min(v, a v): FMUL + FMNMX (the predicate is
// warp-uniform and hoisted), one instruction less than compare + select.  Bit-identical to the select form.
__device__ __forceinline__ float prelu(float v, float a) { const float m = a * v; return a <= 1.f ? fmaxf(v, m) : fminf(v, m); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float f4get(const float4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(std::integral_constant<int, I>{}), ...); }
// compile-time loop: f(integral_constant<int, 0>) .. f(integral_constant<int, N-1>), guaranteed unrolled
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// Packed fp32 FMA (sm_100 FFMA2, fma.rn.f32x2): two IEEE fmas in one issue slot.  The convolutions here are
// acc[c] += x * w[c] with x per thread and w[c] warp-uniform, which maps to `FFMA2 Racc, Rx.F32, URw.F32x2, Racc` (scalar
// x broadcast, weight pair in uniform registers fed by LDCU.128): half the FMA issue slots and half the constant loads of
// the FFMA form, results bit-identical (same fma per lane).  -DGTCRNB_NO_FFMA2 restores the scalar form for A/B runs.
__device__ __forceinline__ void fma_pair(float& a0, float& a1, float x, float w0, float w1) {
#ifndef GTCRNB_NO_FFMA2
    const float2 r = __ffma2_rn(make_float2(x, x), make_float2(w0, w1), make_float2(a0, a1));
    a0 = r.x; a1 = r.y;
#else
    a0 = fmaf(x, w0, a0); a1 = fmaf(x, w1, a1);
#endif
}
// PReLU of two values with one packed multiply (FMUL2) + two FMNMX; bit-identical to prelu() per element
__device__ __forceinline__ void prelu_pair(float& v0, float& v1, float a) {
#ifndef GTCRNB_NO_FFMA2
    const float2 m = __fmul2_rn(make_float2(v0, v1), make_float2(a, a));
    v0 = a <= 1.f ? fmaxf(v0, m.x) : fminf(v0, m.x);
    v1 = a <= 1.f ? fmaxf(v1, m.y) : fminf(v1, m.y);
#else
    v0 = prelu(v0, a); v1 = prelu(v1, a);
#endif
}
__device__ __forceinline__ void add_pair(float& a0, float& a1, float b0, float b1) {
#ifndef GTCRNB_NO_FFMA2
    const float2 r = __fadd2_rn(make_float2(a0, a1), make_float2(b0, b1));
    a0 = r.x; a1 = r.y;
#else
    a0 += b0; a1 += b1;
#endif
}
// a0 += x0 * w0, a1 += x1 * w1 (distinct inputs: depthwise taps, adjacent channels of one float4)
__device__ __forceinline__ void fma_pair2(float& a0, float& a1, float x0, float x1, float w0, float w1) {
#ifndef GTCRNB_NO_FFMA2
    const float2 r = __ffma2_rn(make_float2(x0, x1), make_float2(w0, w1), make_float2(a0, a1));
    a0 = r.x; a1 = r.y;
#else
    a0 = fmaf(x0, w0, a0); a1 = fmaf(x1, w1, a1);
#endif
}

__device__ __forceinline__ void copy_f4(float* dst, const float* __restrict__ src, int n_floats, int tid, int nthr) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = tid; i < (n_floats >> 2); i += nthr) d4[i] = s4[i];
}
// asynchronous variant (cp.async, LDGSTS): global -> shared without staging registers, so the copy overlaps whatever the thread
// does next; complete it with cp_async_wait_all() before the barrier that publishes the data
__device__ __forceinline__ void copy_f4_async(float* dst, const float* __restrict__ src, int n_floats, int tid, int nthr) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    for (int i = tid; i < (n_floats >> 2); i += nthr)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * i), "l"(src + 4 * i) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void zero_f4(float* dst, int n_floats, int tid, int nthr) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = tid; i < (n_floats >> 2); i += nthr) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// acc[0..N) += x * w[0..N)  with w read as broadcast LDS.128 (N multiple of 4)
template <int N>
__device__ __forceinline__ void axpy2(float (&a0)[N], float (&a1)[N], float x0, float x1, const float* __restrict__ w) {
    const float4* w4 = reinterpret_cast<const float4*>(w);
#pragma unroll
    for (int v = 0; v < N / 4; ++v) {
        const float4 ww = w4[v];
        fma_pair(a0[4 * v + 0], a0[4 * v + 1], x0, ww.x, ww.y); fma_pair(a1[4 * v + 0], a1[4 * v + 1], x1, ww.x, ww.y);
        fma_pair(a0[4 * v + 2], a0[4 * v + 3], x0, ww.z, ww.w); fma_pair(a1[4 * v + 2], a1[4 * v + 3], x1, ww.z, ww.w);
    }
}

// same with the weights coming from a kernel parameter struct (constant bank, warp-uniform index): no LDS, no
// staging registers -- ptxas turns the reads into LDCU + FFMA-with-uniform-register
template <int N>
__device__ __forceinline__ void axpy2c(float (&a0)[N], float (&a1)[N], float x0, float x1, const float* w) {
#pragma unroll
    for (int c = 0; c < N; c += 2) {
        const float w0 = w[c], w1 = w[c + 1];
        fma_pair(a0[c], a0[c + 1], x0, w0, w1);
        fma_pair(a1[c], a1[c + 1], x1, w0, w1);
    }
}

#ifndef GTCRNB_DEC_UNROLL
#define GTCRNB_DEC_UNROLL 1      // decoder 3x3 (kf, ci-quad) loop: 0 rolled (12 iterations), 1 kf rolled, 2 fully unrolled
#endif
#ifndef GTCRNB_TCN_UNROLL
#define GTCRNB_TCN_UNROLL 0      // GTCN: 0 loop over the four TCNs rolled (weights at a uniform-register offset), 1 unrolled
#endif
#ifndef GTCRNB_EN1_UNROLL
#define GTCRNB_EN1_UNROLL 1      // front en1 tap loop: 0 rolled, 1 unrolled
#endif

template <int NF> struct alignas(16) ParamBlock { float w[NF]; };
struct SmemWeights { const float* __restrict__ w; };      // same .w[] interface as ParamBlock, backed by shared memory
// compile-time choice of the weight source of a kernel (shared-memory copy or the kernel-parameter block itself)
template <bool SMEM, class S, class C>
__device__ __forceinline__ const auto& pick_weights(const S& smem, const C& cbank) {
    if constexpr (SMEM) return smem; else return cbank;
}

// ================================================================================================
// 256-point complex FFT, radix-4 Stockham autosort, 64 threads per transform, 4 stages.
// (the 512-point real transforms of torch.stft / torch.istft are done as one 256-point complex FFT
//  plus an O(N) split step; reference call sites infer.py:60-67, :73-76)
// ================================================================================================
__device__ __forceinline__ void fft256_stage(const float2* __restrict__ in, float2* __restrict__ out,
                                             const float2* __restrict__ tw, int j, int Ns) {
    const int k = j & (Ns - 1);
    float2 v0 = in[j], v1 = in[j + 64], v2 = in[j + 128], v3 = in[j + 192];
    if (Ns > 1) {
        // stage-major table, three twiddles per butterfly at consecutive float2s (pack.py _fft_tables): conflict-free
        const float2* tws = tw + (Ns == 4 ? 0 : Ns == 16 ? 12 : 60) + 3 * k;
        v1 = cmul(v1, tws[0]);
        v2 = cmul(v2, tws[1]);
        v3 = cmul(v3, tws[2]);
    }
    const float2 s0 = make_float2(v0.x + v2.x, v0.y + v2.y), s1 = make_float2(v0.x - v2.x, v0.y - v2.y);
    const float2 s2 = make_float2(v1.x + v3.x, v1.y + v3.y), d = make_float2(v1.x - v3.x, v1.y - v3.y);
    const float2 s3 = make_float2(d.y, -d.x);   // d * (-i)
    const int j0 = ((j - k) << 2) + k;
    if (Ns == 1) {                                          // the four outputs are contiguous: two 128-bit stores (half the wavefronts)
        float4* o4 = reinterpret_cast<float4*>(out + 4 * j);
        o4[0] = make_float4(s0.x + s2.x, s0.y + s2.y, s1.x + s3.x, s1.y + s3.y);
        o4[1] = make_float4(s0.x - s2.x, s0.y - s2.y, s1.x - s3.x, s1.y - s3.y);
        return;
    }
    out[j0] = make_float2(s0.x + s2.x, s0.y + s2.y);
    out[j0 + Ns] = make_float2(s1.x + s3.x, s1.y + s3.y);
    out[j0 + 2 * Ns] = make_float2(s0.x - s2.x, s0.y - s2.y);
    out[j0 + 3 * Ns] = make_float2(s1.x - s3.x, s1.y - s3.y);
}
// all threads of the block must call this (it contains __syncthreads); result ends in `a`
__device__ __forceinline__ void fft256(float2* a, float2* b, const float2* tw, int j) {
    fft256_stage(a, b, tw, j, 1);  __syncthreads();
    fft256_stage(b, a, tw, j, 4);  __syncthreads();
    fft256_stage(a, b, tw, j, 16); __syncthreads();
    fft256_stage(b, a, tw, j, 64); __syncthreads();
}

// same, for blocks with more than 4 x 64 threads: threads with active == false only take part in the barriers
__device__ __forceinline__ void fft256_masked(float2* a, float2* b, const float2* tw, int j, bool active) {
    if (active) fft256_stage(a, b, tw, j, 1);
    __syncthreads();
    if (active) fft256_stage(b, a, tw, j, 4);
    __syncthreads();
    if (active) fft256_stage(a, b, tw, j, 16);
    __syncthreads();
    if (active) fft256_stage(b, a, tw, j, 64);
    __syncthreads();
}

// one WARP per transform: every lane does two radix-4 butterflies per stage; only __syncwarp between stages, so 8 frames can
// be transformed by 8 warps without a single CTA barrier (the 64-thread version cost k_front 35 % of its samples in barriers)
__device__ __forceinline__ void fft256_warp(float2* a, float2* b, const float2* tw, int lane) {
    fft256_stage(a, b, tw, lane, 1);  fft256_stage(a, b, tw, lane + 32, 1);  __syncwarp();
    fft256_stage(b, a, tw, lane, 4);  fft256_stage(b, a, tw, lane + 32, 4);  __syncwarp();
    fft256_stage(a, b, tw, lane, 16); fft256_stage(a, b, tw, lane + 32, 16); __syncwarp();
    fft256_stage(b, a, tw, lane, 64); fft256_stage(b, a, tw, lane + 32, 64); __syncwarp();
}

__device__ __forceinline__ int reflect_idx(int i, int L) {
    if (i < 0) i = -i;
    if (i >= L) i = 2 * (L - 1) - i;
    return i;
}

struct FftTables {
    float2 tw256[256];
    float2 tw512[258];
    float win[512];
};
__device__ __forceinline__ void load_fft_tables(FftTables& t, const float* __restrict__ blobp, int tid, int nthr) {
    const float* f = blobp + blob::FFT_OFF;
    copy_f4(reinterpret_cast<float*>(t.tw256), f + blob::FFT_TW256, 512, tid, nthr);
    copy_f4(reinterpret_cast<float*>(t.tw512), f + blob::FFT_TW512, 516, tid, nthr);
    copy_f4(t.win, f + blob::FFT_WIN, 512, tid, nthr);
}
__device__ __forceinline__ void load_fft_tables_async(FftTables& t, const float* __restrict__ blobp, int tid, int nthr) {
    const float* f = blobp + blob::FFT_OFF;
    copy_f4_async(reinterpret_cast<float*>(t.tw256), f + blob::FFT_TW256, 512, tid, nthr);
    copy_f4_async(reinterpret_cast<float*>(t.tw512), f + blob::FFT_TW512, 516, tid, nthr);
    copy_f4_async(t.win, f + blob::FFT_WIN, 512, tid, nthr);
}

// ------------------------------------------------------------------------------------------------
// STFT: wav [B][L] -> spec [B][T][257][2], T = 1 + L/256.  centre=True, reflect padding, sqrt-Hann.
// grid (ceil(T/4), B), 256 threads = 4 frames x 64 threads.
// ------------------------------------------------------------------------------------------------
constexpr int STFT_FPB = 4;
__global__ void __launch_bounds__(256) k_stft(const float* __restrict__ wav, int64_t in_stride, int L, int T,
                                              const float* __restrict__ blobp, float* __restrict__ spec) {
    __shared__ FftTables tb;
    __shared__ __align__(16) float2 buf[STFT_FPB][2][256];
    const int tid = threadIdx.x, fr = tid >> 6, j = tid & 63;
    const int t = blockIdx.x * STFT_FPB + fr, b = blockIdx.y;
    const bool valid = t < T;
    load_fft_tables(tb, blobp, tid, 256);
    __syncthreads();
    float2* A = buf[fr][0];
    float2* Bf = buf[fr][1];
    if (valid) {
        const float* x = wav + (int64_t)b * in_stride;
        const int base = kHop * t - kHop;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int n = j + 64 * r;
            const int i0 = reflect_idx(base + 2 * n, L), i1 = reflect_idx(base + 2 * n + 1, L);
            A[n] = make_float2(tb.win[2 * n] * x[i0], tb.win[2 * n + 1] * x[i1]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) A[j + 64 * r] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft256(A, Bf, tb.tw256, j);
    if (valid) {
        float2* o = reinterpret_cast<float2*>(spec + ((int64_t)b * T + t) * kSpecFrame);
        for (int k = j; k <= 256; k += 64) {
            const float2 zk = A[k & 255];
            float2 zc = A[(256 - k) & 255];
            zc.y = -zc.y;
            const float2 w = tb.tw512[k];
            const float dx = zk.x - zc.x, dy = zk.y - zc.y;
            o[k] = make_float2(0.5f * (zk.x + zc.x) + 0.5f * (w.x * dy + w.y * dx),
                               0.5f * (zk.y + zc.y) - 0.5f * (w.x * dx - w.y * dy));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// iSTFT + overlap-add: spec [B][T][257][2] (element strides sb, sf, st, sc) -> wav [B][256 (T-1)].
// grid (ceil((T-1)/ISTFT_HOPS), B), 256 threads.  A block inverts ISTFT_HOPS+1 frames (4 at a time).
// ------------------------------------------------------------------------------------------------
constexpr int ISTFT_HOPS = 7;
__global__ void __launch_bounds__(256) k_istft(const float* __restrict__ spec, int64_t sb, int64_t sf, int64_t st, int64_t sc,
                                               int T, const float* __restrict__ blobp, float* __restrict__ wav, int64_t out_stride) {
    __shared__ FftTables tb;
    __shared__ __align__(16) float2 buf[4][2][256];
    __shared__ float frames[ISTFT_HOPS + 1][512];
    __shared__ float inv_env[256];
    const int tid = threadIdx.x, fr = tid >> 6, j = tid & 63;
    const int b = blockIdx.y, k0 = blockIdx.x * ISTFT_HOPS;
    const int nhop = min(ISTFT_HOPS, (T - 1) - k0);
    load_fft_tables(tb, blobp, tid, 256);
    inv_env[tid] = blobp[blob::FFT_OFF + blob::FFT_INV_ENV + tid];
    __syncthreads();
    const float* sp = spec + (int64_t)b * sb;
    for (int r0 = 0; r0 < nhop + 1; r0 += 4) {
        const int fi = r0 + fr;
        const bool valid = fi < nhop + 1;
        const int t = k0 + fi;
        float2* A = buf[fr][0];
        float2* Bf = buf[fr][1];
        if (valid) {
            const float* xt = sp + (int64_t)t * st;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int k = j + 64 * r;                       // 0..255
                float2 xk = make_float2(xt[(int64_t)k * sf], xt[(int64_t)k * sf + sc]);
                float2 xc = make_float2(xt[(int64_t)(256 - k) * sf], -xt[(int64_t)(256 - k) * sf + sc]);
                if (k == 0) { xk.y = 0.f; xc.y = 0.f; }         // imag of DC / Nyquist ignored (C2R semantics)
                const float2 w = tb.tw512[k];
                const float dx = xk.x - xc.x, dy = xk.y - xc.y;
                const float a = w.x * dx + w.y * dy, bq = w.x * dy - w.y * dx;   // conj(w) * D
                const float zx = 0.5f * (xk.x + xc.x) - 0.5f * bq, zy = 0.5f * (xk.y + xc.y) + 0.5f * a;
                A[k] = make_float2(zx, -zy);                    // conj(Z): inverse FFT through the forward kernel
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) A[j + 64 * r] = make_float2(0.f, 0.f);
        }
        __syncthreads();
        fft256(A, Bf, tb.tw256, j);
        if (valid) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int m = j + 64 * r;
                const float2 z = A[m];
                frames[fi][2 * m] = tb.win[2 * m] * (z.x * (1.0f / 256.0f));
                frames[fi][2 * m + 1] = tb.win[2 * m + 1] * (-z.y * (1.0f / 256.0f));
            }
        }
        __syncthreads();
    }
    float* o = wav + (int64_t)b * out_stride;
    for (int h = 0; h < nhop; ++h)
        o[(int64_t)(k0 + h) * kHop + tid] = (frames[h][256 + tid] + frames[h + 1][tid]) * inv_env[tid];
}

// ================================================================================================
// Tensor-core ends (round 2): en1 in k_front and de3 in k_back -- the two 16 -> 16, 5-tap strided (transposed) convolutions, 70 % of
// those kernels' MACs -- as tcgen05 kind::f16 MMAs with BOTH operands in shared memory.  A frame's activations already sit in
// shared memory as K-major planes [ci chunk][row = slot * 35 + padded bin][8 halves]: rows are 16 bytes apart, so an M = 128 tile
// is 128 consecutive rows and a convolution tap is the same plane read through a descriptor whose start address is shifted by
// whole rows (the stride-2 en1 reads parity-split planes).  Rows that straddle a frame's 33 outputs (pad bins, the tail of the
// third tile) produce accumulator rows nobody reads; a tcgen05 accumulator row depends only on its own row of A.
// Operand format: x = 256 hi' + lo with hi' = fp16(x / 256) and lo = fp16(x - 256 hi') (exact residual, 22 mantissa bits in all;
// the offset of 2^8 between the two halves moves fp16's ceiling to |x| = 1.6e7, far above anything 30 x full-scale audio produces
// -- the conversions saturate instead of overflowing).  Weights: W s split as fp16 hi / lo with a power-of-two s that keeps the
// lo parts normal; the hi' operand meets 256 W s, the lo operand W s, so one accumulator receives s * sum(x W) and the epilogue
// multiplies by 1 / s (exact) while adding the bias.  Three MMAs per tap and 16 input channels: hi' [W_hi | W_lo] stacked along N
// (k_front) or separately (k_back, where N already carries the even | odd output bins), lo W_hi.
// The weight matrices are identical for every CTA: k_prep_ends_tc builds them once per handle in the exact shared-memory layout
// (no-swizzle K-major, [k chunk][n][8 halves]) and the kernels copy them with cp.async.
// ================================================================================================
#ifdef GTCRNB_FB_TRACE
// phase timeline of ONE CTA of k_front / k_back in steady state (clock64 stamps of the first lane of warps 0, 4 and 8): tools/gpu_fb_trace.sh
__device__ long long g_fb_trace[2][3][16];
#define FBT(kern, i) do { if (blockIdx.x == 20000 && (threadIdx.x == 0 || threadIdx.x == 128 || threadIdx.x == 256)) { long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)); g_fb_trace[kern][threadIdx.x >> 7][i] = c_; } } while (0)
#else
#define FBT(kern, i) do { } while (0)
#endif
constexpr int TCW_EN1_WORDS = 5 * 384 + 4;      // 5 taps x ([2][32][8] + [2][16][8] halves) + {1 / s, 0, 0, 0}
constexpr int TCW_DE3_WORDS = 3 * 768 + 4;      // 3 taps x 3 x [2][32][8] halves + {1 / s, 0, 0, 0}
constexpr int TCW_WORDS = TCW_EN1_WORDS + TCW_DE3_WORDS;
constexpr int TCE_ROWS = 8 * 35;                // rows of an activation plane: 8 frame slots x 35 padded bins

__device__ __forceinline__ void split_f16_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = tc::pack_f16x2(x0 * (1.0f / 256.0f), x1 * (1.0f / 256.0f));
    lo = tc::pack_f16x2(fmaf(tc::f16lo_to_f32(hi), -256.0f, x0), fmaf(tc::f16hi_to_f32(hi), -256.0f, x1));
}
// 16 channels of one position -> two uint4 of hi' halves (ci 0..7, 8..15) and two of lo halves
__device__ __forceinline__ void split_f16_row(const float (&v)[16], uint4 (&hi)[2], uint4 (&lo)[2]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) split_f16_pair(v[2 * c], v[2 * c + 1], h[c], l[c]);
    hi[0] = make_uint4(h[0], h[1], h[2], h[3]); hi[1] = make_uint4(h[4], h[5], h[6], h[7]);
    lo[0] = make_uint4(l[0], l[1], l[2], l[3]); lo[1] = make_uint4(l[4], l[5], l[6], l[7]);
}

// one CTA, run once per handle: fp16 hi / lo weight matrices of en1 (blob FRONT_EN1_W [k][ci][co]) and de3 (BACK_DE3_W [k][ci][co])
__global__ void __launch_bounds__(256) k_prep_ends_tc(const float* __restrict__ blobp, uint32_t* __restrict__ tcw) {
    __shared__ unsigned mx[2];
    const int tid = threadIdx.x;
    if (tid < 2) mx[tid] = 0u;
    for (int i = tid; i < TCW_WORDS; i += 256) tcw[i] = 0u;
    __syncthreads();
    const float* w1 = blobp + blob::FRONT_OFF + blob::FRONT_EN1_W;
    const float* w3 = blobp + blob::BACK_OFF + blob::BACK_DE3_W;
    unsigned m1 = 0u, m3 = 0u;
    for (int i = tid; i < 1280; i += 256) { m1 = max(m1, __float_as_uint(fabsf(w1[i]))); m3 = max(m3, __float_as_uint(fabsf(w3[i]))); }
    atomicMax(&mx[0], m1); atomicMax(&mx[1], m3);
    __syncthreads();
    __half* hw = reinterpret_cast<__half*>(tcw);
#pragma unroll
    for (int layer = 0; layer < 2; ++layer) {
        // |w| < 2^(e+1): 256 w s < 2^14 with s = 2^(5-e); the lo parts (2^-11 of that) stay normal down to |w| ~ 2^-13 max|w|
        const int wexp = max(14, min(253, (int)(mx[layer] >> 23)));
        const float sa = __uint_as_float((unsigned)(267 - wexp) << 23), sb = __uint_as_float((unsigned)(259 - wexp) << 23);
        const float winv = __uint_as_float((unsigned)(wexp - 5) << 23);
        for (int i = tid; i < 1280; i += 256) {
            const int k = i >> 8, ci = (i >> 4) & 15, co = i & 15, jc = ci >> 3, cl = ci & 7;
            const float w = layer == 0 ? w1[i] : w3[i];
            const float wa = w * sa;
            const __half ah = __float2half_rn(wa), al = __float2half_rn(wa - __half2float(ah)), bh = __float2half_rn(w * sb);
            if (layer == 0) {
                __half* t = hw + k * 768;                                   // [W_hi | W_lo] stacked: n = co, 16 + co
                t[jc * 256 + co * 8 + cl] = ah;
                t[jc * 256 + (16 + co) * 8 + cl] = al;
                t[512 + jc * 128 + co * 8 + cl] = bh;
            } else {
                // transposed conv in gather form: tap pair j = k / 2 reads input bin m + 1 - j; even output bins use tap 2j,
                // odd ones tap 2j + 1 (none for j = 2): n = (k & 1) * 16 + co
                __half* t = hw + 2 * TCW_EN1_WORDS + (k >> 1) * 1536;
                const int n = (k & 1) * 16 + co;
                t[jc * 256 + n * 8 + cl] = ah;
                t[512 + jc * 256 + n * 8 + cl] = al;
                t[1024 + jc * 256 + n * 8 + cl] = bh;
            }
        }
        if (tid == 0) reinterpret_cast<float*>(tcw)[layer == 0 ? TCW_EN1_WORDS - 4 : TCW_WORDS - 4] = winv;
    }
}

// ================================================================================================
// FRONT: spec -> |.|,re,im -> ERB compress -> SFE-Lite -> en0 (3->16, k5 s2) -> en1 (16->16, k5 s2)
// reference: gtcrn_micro.py:512-518 (features, erb.bm, sfe), :344-364 (en_convs[0..1])
// Frames are independent here, so (b,t) is flattened: grid ceil(B*T/8), 288 threads, 8 frames per CTA
// (the streaming path calls the same kernel with T = 1).  Writes en0 [B][T][65][16], en1 [B][T][33][16].
// ================================================================================================
#ifndef GTCRNB_FR_NF
#define GTCRNB_FR_NF 8
#endif
constexpr int FR_NF = GTCRNB_FR_NF;        // frames per CTA (even)
constexpr int FR_HALF = ((FR_NF / 2) * kF2 + 31) / 32 * 32;   // en1 runs its two output-channel halves on separate warps (warp-uniform weights)
constexpr int FR_THREADS = 2 * FR_HALF;    // 320 at 8 frames (warps 0-4 / 5-9), 192 at 4
constexpr int FR_MIN_BLOCKS = FR_NF >= 8 ? 3 : 5;
static_assert(FR_NF % 2 == 0 && FR_THREADS >= (FR_NF / 2) * kF1 && FR_THREADS / 32 >= FR_NF && 8 * FR_THREADS >= FR_NF * kBins, "k_front thread map");
constexpr int FR_FEAT = 260 * 4;           // per slot: one float4 (mag, re, im, 0) per bin, 257 bins
constexpr int FR_ERB = 64 * 4;             // per slot: one float4 per ERB band (bins 65.. of the compressed axis); low 65 bins stay in feat
// sfe and a0 are stored PARITY-SPLIT along the (padded) bin index c: the stride-2 convolutions read c = 2 f' + k, i.e.
// plane (k & 1), entry f' + (k >> 1) -- consecutive lanes hit consecutive 16-byte entries (no 2-way bank conflict)
constexpr int FR_SFE4 = 2 * 67;            // float4: [parity][67], c = f + 2 in [0,133)
constexpr int FR_SFE = FR_SFE4 * 4;
constexpr int FR_A0_4 = 4 * 2 * 35;        // float4: [q][parity][35], c = f' + 2 in [0,69) -- overlays feat|erb (dead once SFE is done)
constexpr int FR_A0 = FR_A0_4 * 4;
constexpr int FR_W_FLOATS = blob::FRONT_EN0_W;     // only the ERB / SFE tables go to shared memory; conv weights are kernel params
constexpr int FR_SLOT = FR_FEAT + FR_ERB + FR_SFE;
static_assert(FR_A0 <= FR_FEAT + FR_ERB, "a0 must fit in the dead feat|erb region");
constexpr int FR_FFT_FLOATS = 1540;         // FftTables: tw256 (512) + tw512 (516) + win (512)
static_assert(sizeof(FftTables) == FR_FFT_FLOATS * 4, "FftTables layout");
// tensor-core en1 (TCE): the slot regions are regrouped as [feat x NF | erb x NF | sfe x NF] so that the dead feat | erb area holds the
// fp16 activation planes of all slots contiguously; plane stride = 282 rows: the two parity planes then sit 64 bytes (mod 128) apart
// and the en0 epilogue's 16-byte stores of neighbouring (odd / even) bins do not collide
constexpr int TCE_PLANE = 282 * 16;          // bytes per activation plane in k_front: [(hl * 2 + parity) * 2 + ci chunk]
static_assert(7 * TCE_PLANE + (3 * 128 + 2) * 16 <= FR_NF * (FR_FEAT + FR_ERB) * 4, "over-read of the last tile stays inside the dead region");
__host__ __device__ constexpr int fr_smem_floats(bool from_wav, bool tce = false) {
    return FR_W_FLOATS + (tce ? TCW_EN1_WORDS : 0) + FR_NF * FR_SLOT + (from_wav ? FR_FFT_FLOATS : 0);
}

// Model variant of the per-frame front / back kernels.  VAR 0: GTCRN-Micro (SFE-Lite, dense en1 / de3).  VAR 1: original-GTCRN-style
// (BASELINE configs[3]): unfold SFE -- en0 contracts 9 channels x 5 taps (gtcrn_micro.py:346 `# 3 * 3,  # SFE`) -- and en1 / de3 with
// groups = 2 (gtcrn_micro.py:361,448 `# groups=2`).  The ERB / SFE tables sit at the same offsets in both blob layouts.
template <int VAR> struct FrontL;
template <> struct FrontL<0> {
    static constexpr int SIZE = blob::FRONT_SIZE, EN0_W = blob::FRONT_EN0_W, EN0_B = blob::FRONT_EN0_B, EN1_W = blob::FRONT_EN1_W,
                         EN1_B = blob::FRONT_EN1_B, SLOPE = blob::FRONT_SLOPE, OFF = blob::FRONT_OFF, FFT = blob::FFT_OFF;
};
template <> struct FrontL<1> {
    static constexpr int SIZE = blobf::FRONT_SIZE, EN0_W = blobf::FRONT_EN0_W, EN0_B = blobf::FRONT_EN0_B, EN1_W = blobf::FRONT_EN1_W,
                         EN1_B = blobf::FRONT_EN1_B, SLOPE = blobf::FRONT_SLOPE, OFF = blobf::FRONT_OFF, FFT = blobf::FFT_OFF;
};
static_assert(blob::FRONT_EN0_W == blobf::FRONT_EN0_W && blob::FRONT_ERB_W == blobf::FRONT_ERB_W && blob::FRONT_SFE_W == blobf::FRONT_SFE_W &&
              blob::FFT_OFF == blobf::FFT_OFF && blob::FFT_SIZE == blobf::FFT_SIZE, "front tables share their offsets between the variants");
template <int VAR> struct BackL;
template <> struct BackL<0> {
    static constexpr int SIZE = blob::BACK_SIZE, OFF = blob::BACK_OFF, DE3_W = blob::BACK_DE3_W, DE3_B = blob::BACK_DE3_B, DE4_W = blob::BACK_DE4_W,
                         DE4_B = blob::BACK_DE4_B, IERB_B0 = blob::BACK_IERB_B0, IERB_W = blob::BACK_IERB_W, SLOPE = blob::BACK_SLOPE;
};
template <> struct BackL<1> {
    static constexpr int SIZE = blobf::BACK_SIZE, OFF = blobf::BACK_OFF, DE3_W = blobf::BACK_DE3_W, DE3_B = blobf::BACK_DE3_B, DE4_W = blobf::BACK_DE4_W,
                         DE4_B = blobf::BACK_DE4_B, IERB_B0 = blobf::BACK_IERB_B0, IERB_W = blobf::BACK_IERB_W, SLOPE = blobf::BACK_SLOPE;
};

// en1 for one output-channel half H (compile-time so the weight indices stay warp-uniform): thread = 2 slots x 8 co at one f''
// VAR 1 (groups = 2): half H only sees input channels 8H .. 8H+7 (ci quads 2H, 2H+1), weights [k][ci][co within the group]
template <int H, int VAR, class PB>
__device__ __forceinline__ void front_en1(const PB& prm, const float* slots, int r, int64_t fi0, int64_t nframes, float* __restrict__ en1) {
    using L = FrontL<VAR>;
    const int sp2 = r / kF2, f = r - sp2 * kF2, s0 = 2 * sp2, s1 = s0 + 1;
    float acc0[8], acc1[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc0[c] = acc1[c] = prm.w[L::EN1_B + 8 * H + c];
    const float4* i0p = reinterpret_cast<const float4*>(slots + s0 * FR_SLOT) + f;
    const float4* i1p = reinterpret_cast<const float4*>(slots + s1 * FR_SLOT) + f;
#if GTCRNB_EN1_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int k = 0; k < 5; ++k) {
        const int po = (k & 1) * 35 + (k >> 1);
#pragma unroll
        for (int q = (VAR == 1 ? 2 * H : 0); q < (VAR == 1 ? 2 * H + 2 : 4); ++q) {
            const float4 x0 = i0p[q * 70 + po], x1 = i1p[q * 70 + po];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const int wi = VAR == 1 ? L::EN1_W + (k * 16 + 4 * q + i) * 8 + c : L::EN1_W + (k * 16 + 4 * q + i) * 16 + 8 * H + c;
                    const float w0 = prm.w[wi];
                    const float w1 = prm.w[wi + 1];
                    fma_pair(acc0[c], acc0[c + 1], f4get(x0, i), w0, w1);
                    fma_pair(acc1[c], acc1[c + 1], f4get(x1, i), w0, w1);
                }
        }
    }
    const float a = prm.w[L::SLOPE + 1];
#pragma unroll
    for (int c = 0; c < 8; c += 2) { prelu_pair(acc0[c], acc0[c + 1], a); prelu_pair(acc1[c], acc1[c + 1], a); }
    if (fi0 + s0 < nframes) {
        float4* g = reinterpret_cast<float4*>(en1 + ((fi0 + s0) * kF2 + f) * 16 + 8 * H);
        g[0] = make_float4(acc0[0], acc0[1], acc0[2], acc0[3]);
        g[1] = make_float4(acc0[4], acc0[5], acc0[6], acc0[7]);
    }
    if (fi0 + s1 < nframes) {
        float4* g = reinterpret_cast<float4*>(en1 + ((fi0 + s1) * kF2 + f) * 16 + 8 * H);
        g[0] = make_float4(acc1[0], acc1[1], acc1[2], acc1[3]);
        g[1] = make_float4(acc1[4], acc1[5], acc1[6], acc1[7]);
    }
}

// FROM_WAV = false: `spec` is the input spectrum (element strides sb, sf, st, sc), `wav`/`spec_out` unused.
// FROM_WAV = true : the STFT is fused in: frames are cut from `wav` [B][L] (reflect padding, sqrt-Hann), transformed in
//                   shared memory (the FFT ping-pong buffers overlay the slot's own feat / sfe regions) and the spectrum is
//                   written once to `spec_out` [B][T][257][2] for the mask stage; it is never read back here.
template <bool FROM_WAV, int VAR, bool TCE = false>
__global__ void __launch_bounds__(FR_THREADS, FR_MIN_BLOCKS) k_front(const float* __restrict__ spec, int64_t sb, int64_t sf, int64_t st, int64_t sc,
                                                         const float* __restrict__ wav, int64_t wav_stride, int L, float* __restrict__ spec_out,
                                                         int T, int64_t nframes, const float* __restrict__ blobp,
                                                         const __grid_constant__ ParamBlock<FrontL<VAR>::SIZE> prm,
                                                         float* __restrict__ en0, float* __restrict__ en1, const uint32_t* __restrict__ tcw = nullptr) {
    static_assert(!TCE || (VAR == 0 && FR_NF == 8), "tensor-core en1: GTCRN-Micro, 8 frame slots");
    using LY = FrontL<VAR>;
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) uint64_t tc_bars[3];
    __shared__ uint32_t tmem_base_s;
    float* W = reinterpret_cast<float*>(smem4);
    float* Btc = W + FR_W_FLOATS;              // TCE: en1 weight matrices (k_prep_ends_tc layout)
    float* slots = Btc + (TCE ? TCW_EN1_WORDS : 0);   // per slot: [feat | erb | sfe]; a0 overlays feat|erb (TCE: regrouped, see featp)
    const int tid = threadIdx.x;
    const int64_t fi0 = (int64_t)blockIdx.x * FR_NF;              // first flattened frame (b * T + t) of this CTA
    FBT(0, 0);
    auto featp = [&](int s_) { return TCE ? slots + s_ * FR_FEAT : slots + s_ * FR_SLOT; };
    auto erbp = [&](int s_) { return TCE ? slots + FR_NF * FR_FEAT + s_ * FR_ERB : slots + s_ * FR_SLOT + FR_FEAT; };
    auto sfep = [&](int s_) { return TCE ? slots + FR_NF * (FR_FEAT + FR_ERB) + s_ * FR_SFE : slots + s_ * FR_SLOT + FR_FEAT + FR_ERB; };
    // tables (FFT twiddles / window, ERB / SFE tables, en1 weight matrices): identical for every CTA, moved by ONE thread with 1-D TMA
    // bulk copies (was 1,100 cp.async requests per CTA; the clock64 trace showed 5,700 of a CTA's 34,000 cycles waiting for them)
    __shared__ __align__(8) uint64_t tb_bar;
    if (tid == 64) {
        tc::mbar_init(&tb_bar, 1);
        tc::fence_barrier_init();
        tc::mbar_expect_tx(&tb_bar, FR_W_FLOATS * 4 + (TCE ? TCW_EN1_WORDS * 4 : 0) + (FROM_WAV ? FR_FFT_FLOATS * 4 : 0));
        if (FROM_WAV) {
            FftTables& tb0 = *reinterpret_cast<FftTables*>(slots + FR_NF * FR_SLOT);
            const float* fb = blobp + LY::FFT;
            tc::bulk_g2s(tb0.tw256, fb + blob::FFT_TW256, 2048, &tb_bar);
            tc::bulk_g2s(tb0.tw512, fb + blob::FFT_TW512, 2064, &tb_bar);
            tc::bulk_g2s(tb0.win, fb + blob::FFT_WIN, 2048, &tb_bar);
        }
        tc::bulk_g2s(W, blobp + LY::OFF, FR_W_FLOATS * 4, &tb_bar);
        if (TCE) tc::bulk_g2s(Btc, tcw, TCW_EN1_WORDS * 4, &tb_bar);
    }
    static_assert(FR_W_FLOATS % 4 == 0 && TCW_EN1_WORDS % 4 == 0 && FR_FFT_FLOATS * 4 == 2048 + 2064 + 2048, "bulk copy sizes are multiples of 16 bytes");
    if (TCE) {
        if (tid < 32) { tc::tmem_alloc(&tmem_base_s, 128); tc::fence_before(); }     // 3 tiles x 32 accumulator columns
        if (tid == 32) { for (int i = 0; i < 3; ++i) tc::mbar_init(&tc_bars[i], 1); tc::fence_async_smem(); }
    }

    if (FROM_WAV) {
        // ---- a (fused STFT): 4 frames per round, 64 threads per 256-point FFT.  A = slot's sfe region, B = slot's feat region.
        FftTables& tb = *reinterpret_cast<FftTables*>(slots + FR_NF * FR_SLOT);
        // warp w owns frame slot w.  Its 512 samples are requested first, so their latency hides behind the table copy and the
        // barrier (the loads were 8 % of the kernel's samples as long_scoreboard and made the FFT warps finish unevenly: 15 % of
        // all samples sat in the barrier after this phase, ncu r1_v13)
        const int warp = tid >> 5, lane = tid & 31;
        const int64_t fi_w = fi0 + warp;
        const bool valid = warp < FR_NF && fi_w < nframes;
        float xr0[8], xr1[8];
        if (valid) {
            const int64_t b = nframes < ((int64_t)1 << 31) ? (int64_t)((uint32_t)fi_w / (uint32_t)T) : fi_w / T;   // (a 64-bit division is ~100 instructions)
            const int t = (int)(fi_w - b * T);
            const float* x = wav + b * wav_stride;
            const int base = kHop * t - kHop;
            if (base >= 0 && base + kNfft <= L && ((reinterpret_cast<uintptr_t>(x + base) & 7) == 0)) {
                // interior frame: no reflection, sample pairs are aligned 8-byte loads (8 instead of 16 requests per lane)
                const float2* x2 = reinterpret_cast<const float2*>(x + base);
#pragma unroll
                for (int q = 0; q < 8; ++q) { const float2 v = x2[lane + 32 * q]; xr0[q] = v.x; xr1[q] = v.y; }
            } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n = lane + 32 * q;
                xr0[q] = x[reflect_idx(base + 2 * n, L)];
                xr1[q] = x[reflect_idx(base + 2 * n + 1, L)];
            }
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) xr0[q] = xr1[q] = 0.f;
        }
    FBT(0, 1);
        __syncthreads();                                           // the tables' mbarrier is initialised for everybody
        tc::mbar_wait(&tb_bar, 0);
    FBT(0, 2);
        if (warp < FR_NF) {                                        // (only __syncwarp inside)
            const int slot = warp;
            const int64_t fi = fi_w;
            float2* A = reinterpret_cast<float2*>(sfep(slot));
            float2* Bf = reinterpret_cast<float2*>(featp(slot));
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int n = lane + 32 * q;
                A[n] = make_float2(tb.win[2 * n] * xr0[q], tb.win[2 * n + 1] * xr1[q]);
            }
            __syncwarp();
            fft256_warp(A, Bf, tb.tw256, lane);
            float4* fs = reinterpret_cast<float4*>(featp(slot));              // feat region (B is dead now)
            float2* o = valid ? reinterpret_cast<float2*>(spec_out + fi * kSpecFrame) : nullptr;
            for (int k = lane; k <= 256; k += 32) {
                const float2 zk = A[k & 255];
                float2 zc = A[(256 - k) & 255];
                zc.y = -zc.y;
                const float2 w = tb.tw512[k];
                const float dx = zk.x - zc.x, dy = zk.y - zc.y;
                const float re = 0.5f * (zk.x + zc.x) + 0.5f * (w.x * dy + w.y * dx);
                const float im = 0.5f * (zk.y + zc.y) - 0.5f * (w.x * dx - w.y * dy);
                if (valid) o[k] = make_float2(re, im);
                fs[k] = make_float4(sqrtf(re * re + im * im + 1e-12f), re, im, 0.f);
            }
        }
#ifndef GTCRNB_NO_AHEAD_PREFETCH
        else {
            // the two warps without a frame to transform pull the samples of the CTA that will take over this slot (kAheadCtas CTAs
            // ahead, see k_back) from HBM into L2: 8 frames = 9 hops = 9 KB = 72 lines.  [Issued by all threads at kernel start
            // this made the kernel 2 % slower; here it is free.]
            const int64_t fa = fi0 + (int64_t)(3 * 148) * FR_NF;
            const int i = tid - FR_NF * 32;
            if (fa < nframes && nframes < (int64_t)1 << 31) {
                const int ba = (int)fa / T, ta = (int)fa - ba * T;
                for (int line = i; line < 74; line += FR_THREADS - FR_NF * 32) {
                    const int64_t s_first = (int64_t)kHop * ta - kHop + (int64_t)line * 32;
                    if (s_first >= 0 && s_first < L) asm volatile("prefetch.global.L2 [%0];" ::"l"(wav + (int64_t)ba * wav_stride + s_first));
                }
            }
        }
#endif
    FBT(0, 3);
        __syncthreads();
    FBT(0, 4);
        for (int i = tid; i < FR_NF * 4; i += FR_THREADS) {       // sfe pad columns (its region held FFT buffer A until now)
            float4* sfe4 = reinterpret_cast<float4*>(sfep(i >> 2));
            const int c = i & 3;                                  // padded bins 0, 1, 131, 132 -> (parity, entry)
            sfe4[c == 0 ? 0 : c == 1 ? 67 : c == 2 ? 67 + 65 : 66] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else
    // ---- a: features.  All global loads of the tile are issued first (8 independent loads per thread in flight).
    {
        const bool vec2 = sc == 1 && (((sb | st | sf) & 1) == 0) && ((reinterpret_cast<uintptr_t>(spec) & 7) == 0);
        float re[8], im[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * FR_THREADS;
            const int slot = idx / kBins, k = idx - slot * kBins;
            const int64_t fi = fi0 + slot;
            re[i] = 0.f; im[i] = 0.f;
            if (idx < FR_NF * kBins && fi < nframes) {
                const int64_t b = fi / T, t = fi - b * T;
                const float* p = spec + b * sb + t * st + (int64_t)k * sf;
                if (vec2) { const float2 v = *reinterpret_cast<const float2*>(p); re[i] = v.x; im[i] = v.y; }
                else { re[i] = p[0]; im[i] = p[sc]; }
            }
        }
        for (int i = tid; i < FR_NF * 4; i += FR_THREADS) {
            float4* sfe4 = reinterpret_cast<float4*>(sfep(i >> 2));
            const int c = i & 3;                                  // padded bins 0, 1, 131, 132 -> (parity, entry)
            sfe4[c == 0 ? 0 : c == 1 ? 67 : c == 2 ? 67 + 65 : 66] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * FR_THREADS;
            if (idx < FR_NF * kBins) {
                const int slot = idx / kBins, k = idx - slot * kBins;
                reinterpret_cast<float4*>(featp(slot))[k] = make_float4(sqrtf(re[i] * re[i] + im[i] * im[i] + 1e-12f), re[i], im[i], 0.f);
            }
        }
    }
    FBT(0, 5);
    __syncthreads();
    if (!FROM_WAV) tc::mbar_wait(&tb_bar, 0);                      // ERB / SFE tables (and the en1 matrices) have landed
    FBT(0, 6);
    // ---- b: ERB compress (sparse triangular bank, <= 12 contiguous taps per band): thread = (band, slot pair); each tap weight
    //      is loaded once and applied to the (mag, re, im) float4 of two slots.  The 65 low bins are used in place (no copy).
    if (tid < 64 * (FR_NF / 2)) {
        const int bnd = tid & 63, sp2 = tid >> 6;
        const int s = reinterpret_cast<const int*>(W + blob::FRONT_ERB_START)[bnd], n = reinterpret_cast<const int*>(W + blob::FRONT_ERB_CNT)[bnd];
        const float* ew = W + blob::FRONT_ERB_W + bnd * blob::kErbMaxTap;
        const float4* f0 = reinterpret_cast<const float4*>(featp(2 * sp2)) + 65 + s;
        const float4* f1 = reinterpret_cast<const float4*>(featp(2 * sp2 + 1)) + 65 + s;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        for (int i = 0; i < n; ++i) {
            const float w = ew[i];
            const float4 x0 = f0[i], x1 = f1[i];
            a0.x = fmaf(w, x0.x, a0.x); a0.y = fmaf(w, x0.y, a0.y); a0.z = fmaf(w, x0.z, a0.z);
            a1.x = fmaf(w, x1.x, a1.x); a1.y = fmaf(w, x1.y, a1.y); a1.z = fmaf(w, x1.z, a1.z);
        }
        reinterpret_cast<float4*>(erbp(2 * sp2))[bnd] = a0;
        reinterpret_cast<float4*>(erbp(2 * sp2 + 1))[bnd] = a1;
    }
    FBT(0, 7);
    __syncthreads();
    FBT(0, 8);
    // ---- c: SFE-Lite (depthwise 3-tap along the 129 compressed bins, zero padded, no bias)
    {
        float sw[9];                                               // [c][k]: in registers (the stores below would otherwise force a re-load per item)
#pragma unroll
        for (int i = 0; i < 9; ++i) sw[i] = W[blob::FRONT_SFE_W + i];
        for (int idx = tid; idx < FR_NF * kErb; idx += FR_THREADS) {
            const int slot = idx / kErb, f = idx - slot * kErb;
            const float4* lo = reinterpret_cast<const float4*>(featp(slot));   // compressed bins 0..64 = spectrum bins
            const float4* hi = reinterpret_cast<const float4*>(erbp(slot));    // compressed bins 65..128 = ERB bands
            float4 t[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int j = f + k - 1;
                t[k] = (j < 0 || j > 128) ? make_float4(0.f, 0.f, 0.f, 0.f) : (j < 65 ? lo[j] : hi[j - 65]);
            }
            float4 v;
            v.x = sw[0] * t[0].x + sw[1] * t[1].x + sw[2] * t[2].x;
            v.y = sw[3] * t[0].y + sw[4] * t[1].y + sw[5] * t[2].y;
            v.z = sw[6] * t[0].z + sw[7] * t[1].z + sw[8] * t[2].z;
            v.w = 0.f;
            const int c = f + 2;
            reinterpret_cast<float4*>(sfep(slot))[(c & 1) * 67 + (c >> 1)] = v;
        }
    }
    FBT(0, 9);
    __syncthreads();
    FBT(0, 10);
    // ---- d: en0  (thread = 2 slots x 16 co at one f').  Output a0 overlays the dead feat|erb region; its pad bins are zeroed here.
    if (TCE) {
        // pad rows of the fp16 planes: entries 0 and 34 of both parities and entry 33 of the odd one (padded bins 0, 1, 67, 68 and
        // the unused 69), for 8 slots x 8 planes: 320 16-byte rows, zeroed by the 60 threads that have no en0 position
        if (tid >= (FR_NF / 2) * kF1) {
            uint4* pl = reinterpret_cast<uint4*>(slots);
            for (int i = tid - (FR_NF / 2) * kF1; i < FR_NF * 8 * 5; i += FR_THREADS - (FR_NF / 2) * kF1) {
                const int slot = i / 40, r = i - slot * 40, plane = r / 5, c = r - plane * 5;       // plane = (hl * 2 + parity) * 2 + chunk
                const int par = (plane >> 1) & 1, e = c == 0 ? 0 : c == 1 ? 34 : c == 2 ? 33 : -1;
                if (e >= 0 && (par == 1 || c < 2)) pl[plane * (TCE_PLANE / 16) + slot * 35 + e] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
    } else if (tid >= FR_HALF && tid < FR_HALF + FR_NF * 16) {
        const int i = tid - FR_HALF, slot = i >> 4, q = (i >> 2) & 3, c = i & 3;   // padded bins 0, 1, 67, 68
        reinterpret_cast<float4*>(featp(slot))[q * 70 + (c == 0 ? 0 : c == 1 ? 35 : c == 2 ? 35 + 33 : 34)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (tid < (FR_NF / 2) * kF1) {
        const int sp2 = tid / kF1, f = tid - sp2 * kF1, s0 = 2 * sp2, s1 = s0 + 1;
        float acc0[16], acc1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc0[c] = acc1[c] = prm.w[LY::EN0_B + c];
        const float4* i0p = reinterpret_cast<const float4*>(sfep(s0)) + f;
        const float4* i1p = reinterpret_cast<const float4*>(sfep(s1)) + f;
        if constexpr (VAR == 0) {
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float4 x0 = i0p[(k & 1) * 67 + (k >> 1)], x1 = i1p[(k & 1) * 67 + (k >> 1)];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float w0 = prm.w[LY::EN0_W + k * 48 + i * 16 + c], w1 = prm.w[LY::EN0_W + k * 48 + i * 16 + c + 1];
                        fma_pair(acc0[c], acc0[c + 1], f4get(x0, i), w0, w1);
                        fma_pair(acc1[c], acc1[c + 1], f4get(x1, i), w0, w1);
                    }
            }
        } else {
            // unfold SFE + en0: the conv sees 9 channels (c*3 + k', k' = neighbour bin p-1, p, p+1) at the unfolded positions
            // p = 2f' + k - 2; positions outside [0,128] are the conv's own zero padding (whole tap dropped), neighbours outside
            // are the unfold's zero padding (zero columns of the feature row).  Padded bin of (k, k') = 2f' + (k + k') + 1 - 2:
            // the 7 distinct float4 (j = k + k') are loaded once.
            float4 in0[7], in1[7];
            const float4* b0p = i0p - f;
            const float4* b1p = i1p - f;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int c = 2 * f + j - 1;                                    // padded bin (= bin + 2) of tap pair k + k' = j
                const bool ok = c >= 0 && c <= 132;
                const int e = ok ? (c & 1) * 67 + (c >> 1) : 0;
                in0[j] = ok ? b0p[e] : make_float4(0.f, 0.f, 0.f, 0.f);
                in1[j] = ok ? b1p[e] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int p = 2 * f + k - 2;
                const bool pv = p >= 0 && p <= 128;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
                    float4 x0 = in0[k + kk], x1 = in1[k + kk];
                    if (!pv) { x0 = make_float4(0.f, 0.f, 0.f, 0.f); x1 = x0; }
#pragma unroll
                    for (int i = 0; i < 3; ++i)
#pragma unroll
                        for (int c = 0; c < 16; c += 2) {
                            const float w0 = prm.w[LY::EN0_W + ((k * 3 + kk) * 3 + i) * 16 + c], w1 = prm.w[LY::EN0_W + ((k * 3 + kk) * 3 + i) * 16 + c + 1];
                            fma_pair(acc0[c], acc0[c + 1], f4get(x0, i), w0, w1);
                            fma_pair(acc1[c], acc1[c + 1], f4get(x1, i), w0, w1);
                        }
                }
            }
        }
        const float a = prm.w[LY::SLOPE + 0];
#pragma unroll
        for (int c = 0; c < 16; c += 2) { prelu_pair(acc0[c], acc0[c + 1], a); prelu_pair(acc1[c], acc1[c + 1], a); }
        const int cp = f + 2, po = (cp & 1) * 35 + (cp >> 1);
        if (TCE) {
            // A operand of the en1 MMAs: fp16 hi' / lo rows of the parity plane of this bin (row = slot * 35 + entry)
            uint4* pl = reinterpret_cast<uint4*>(slots) + (cp & 1) * 2 * (TCE_PLANE / 16) + (cp >> 1);
            uint4 hi[2], lo[2];
            split_f16_row(acc0, hi, lo);
            pl[s0 * 35] = hi[0]; pl[(TCE_PLANE / 16) + s0 * 35] = hi[1];
            pl[4 * (TCE_PLANE / 16) + s0 * 35] = lo[0]; pl[5 * (TCE_PLANE / 16) + s0 * 35] = lo[1];
            split_f16_row(acc1, hi, lo);
            pl[s1 * 35] = hi[0]; pl[(TCE_PLANE / 16) + s1 * 35] = hi[1];
            pl[4 * (TCE_PLANE / 16) + s1 * 35] = lo[0]; pl[5 * (TCE_PLANE / 16) + s1 * 35] = lo[1];
        } else {
            float4* o0 = reinterpret_cast<float4*>(featp(s0));
            float4* o1 = reinterpret_cast<float4*>(featp(s1));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                o0[q * 70 + po] = make_float4(acc0[4 * q], acc0[4 * q + 1], acc0[4 * q + 2], acc0[4 * q + 3]);
                o1[q * 70 + po] = make_float4(acc1[4 * q], acc1[4 * q + 1], acc1[4 * q + 2], acc1[4 * q + 3]);
            }
        }
        if (fi0 + s0 < nframes) {
            float4* g = reinterpret_cast<float4*>(en0 + ((fi0 + s0) * kF1 + f) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) g[q] = make_float4(acc0[4 * q], acc0[4 * q + 1], acc0[4 * q + 2], acc0[4 * q + 3]);
        }
        if (fi0 + s1 < nframes) {
            float4* g = reinterpret_cast<float4*>(en0 + ((fi0 + s1) * kF1 + f) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) g[q] = make_float4(acc1[4 * q], acc1[4 * q + 1], acc1[4 * q + 2], acc1[4 * q + 3]);
        }
    }
    if constexpr (TCE) {
        // ---- e (tensor cores): en1 = 5 taps x (hi' [W_hi | W_lo], lo W_hi) per 128-row tile; tap k reads parity plane k & 1 shifted
        //      by k >> 1 rows.  Accumulator of tile i: TMEM columns 32 i .. 32 i + 31 (hi' W_hi + lo W_hi | hi' W_lo).
        tc::fence_async_smem();                                    // the planes (generic-proxy stores) and the cp.async'ed weights
    FBT(0, 11);
        __syncthreads();
    FBT(0, 12);
        tc::fence_after();
        const uint32_t tm = tmem_base_s;
        if (tid == 0) {
            constexpr uint32_t id32 = tc::make_idesc_f16(128, 32), id16 = tc::make_idesc_f16(128, 16);
            const uint32_t a_base = tc::smem_u32(slots), b_base = tc::smem_u32(Btc);
#pragma unroll 1
            for (int tile = 0; tile < 3; ++tile) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const uint32_t a_hi = a_base + (uint32_t)((k & 1) * 2 * TCE_PLANE + (128 * tile + (k >> 1)) * 16);
                    const uint64_t da_hi = tc::make_desc(a_hi, TCE_PLANE, 128), da_lo = tc::make_desc(a_hi + 4 * TCE_PLANE, TCE_PLANE, 128);
                    const uint64_t db_st = tc::make_desc(b_base + k * 1536, 512, 128), db_lo = tc::make_desc(b_base + k * 1536 + 1024, 256, 128);
                    tc::mma_f16_ss(tm + 32 * tile, da_hi, db_st, id32, k > 0);
                    tc::mma_f16_ss(tm + 32 * tile, da_lo, db_lo, id16, 1);
                }
                tc::commit(&tc_bars[tile]);
            }
        }
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < 9) {
            // accumulator rows back: warps 0-3 tile 0, 4-7 tile 1, warp 8 the 24 live rows of tile 2 (a warp reads TMEM lanes 32 (w % 4) ..)
            const int tile = warp < 8 ? warp >> 2 : 2, lg = warp & 3;
            tc::mbar_wait(&tc_bars[tile], 0);
            tc::fence_after();
            float d0[16], d1[16];
            const uint32_t ta = tm + ((uint32_t)(lg * 32) << 16) + 32 * tile;
            tc::tmem_ld16_nowait(ta, d0);
            tc::tmem_ld16_nowait(ta + 16, d1);
            tc::tmem_wait_ld(d0);
            tc::tmem_wait_ld(d1);
            const int r = 128 * tile + 32 * lg + lane, slot = r / 35, e = r - slot * 35;
            const float winv = Btc[TCW_EN1_WORDS - 4], a = prm.w[LY::SLOPE + 1];
#pragma unroll
            for (int c = 0; c < 16; c += 2) {
                add_pair(d0[c], d0[c + 1], d1[c], d1[c + 1]);
                d0[c] = fmaf(d0[c], winv, prm.w[LY::EN1_B + c]); d0[c + 1] = fmaf(d0[c + 1], winv, prm.w[LY::EN1_B + c + 1]);
                prelu_pair(d0[c], d0[c + 1], a);
            }
            if (r < TCE_ROWS && e < kF2 && fi0 + slot < nframes) {
                float4* g = reinterpret_cast<float4*>(en1 + ((fi0 + slot) * kF2 + e) * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) g[q] = make_float4(d0[4 * q], d0[4 * q + 1], d0[4 * q + 2], d0[4 * q + 3]);
            }
        }
        tc::fence_before();
    FBT(0, 13);
        __syncthreads();
    FBT(0, 14);
        if (tid < 32) tc::tmem_dealloc(tm, 128);
    } else {
    __syncthreads();
    // ---- e: en1: output-channel half 0 on warps 0-4, half 1 on warps 5-9 (132 of 160 lanes each)
    {
        const int half = tid >= FR_HALF, r = tid - half * FR_HALF;
        if (r < (FR_NF / 2) * kF2) {
            if (half == 0) front_en1<0, VAR>(prm, slots, r, fi0, nframes, en1);
            else front_en1<1, VAR>(prm, slots, r, fi0, nframes, en1);
        }
    }
    }
}

// ================================================================================================
// GTConvBlock (+TRALite +channel shuffle), encoder (depthwise 3x3) and decoder (full 16->16 3x3) variants.
// reference: gtcrn_micro.py:167-253, TRALite :122-139, shuffle :222-227, decoder skip-add :464-469
//
// One CTA walks one utterance segment [s0, s1) in tiles of NT = 8*PS frames.  A thread owns PS consecutive frame
// slots x 16 channels at one bin, so the PS+2 tap rows it reads per (kf, ci-quad) are shared by its PS outputs.
// Weights are a kernel parameter (constant bank): every weight is one LDCU feeding PS FFMAs, no LDS, no staging
// registers (the v1 kernel spent half its shared-memory wavefronts on broadcast weight loads, ncu r1_v1).
// The point-conv-1 rows live in a RING of NT+2 rows, so the 2-frame causal history is simply the previous rows of
// the ring (no carry copy), and the TRALite gate is evaluated per thread: 3 barriers per tile.
// Segments that do not start at t = 0 warm up on the 4 preceding frames (exactly the dependency cone).
// grid (nseg, B), G*33 active threads (G slot groups x 33 bins), rounded up to whole warps.
// ================================================================================================
constexpr int GT_ROW4 = 4 * 35;                 // float4 per ring row: [ci quad][col = f + 1], cols 0 and 34 stay zero
constexpr int GT_WARMUP = 4;
template <bool DEC> __host__ __device__ constexpr int gt_chunk() { return DEC ? blob::DECGT_SIZE : blob::ENCGT_SIZE; }
// PS frame slots per thread, G slot groups per CTA: NT = G*PS frames per tile, G*33 active threads
__host__ __device__ constexpr int gt_threads(int G) { return (G * kF2 + 31) / 32 * 32; }
__host__ __device__ constexpr int gt_smem_floats(int PS, int G, int w_floats = 0) {
    return (G * PS + 2) * GT_ROW4 * 4 + G * PS * kF2 * 8 + (G * PS + 2) * 8 + G * PS * 8 + w_floats;
}
// CTAs per SM the register budget should allow: PS frame slots per thread need about 56 / 72 / 96 / 128 registers
__host__ __device__ constexpr int gt_min_blocks(int PS, int G = 8) {
    return 65536 / (gt_threads(G) * (PS >= 4 ? 128 : PS == 3 ? 96 : PS == 2 ? 72 : 56));
}

template <bool DEC> struct GtOff {
    static constexpr int PC1_W = DEC ? blob::DECGT_PC1_W : blob::ENCGT_PC1_W;
    static constexpr int PC1_B = DEC ? blob::DECGT_PC1_B : blob::ENCGT_PC1_B;
    static constexpr int MID_W = DEC ? blob::DECGT_FULL_W : blob::ENCGT_DW_W;
    static constexpr int MID_B = DEC ? blob::DECGT_FULL_B : blob::ENCGT_DW_B;
    static constexpr int PC2_W = DEC ? blob::DECGT_PC2_W : blob::ENCGT_PC2_W;
    static constexpr int PC2_B = DEC ? blob::DECGT_PC2_B : blob::ENCGT_PC2_B;
    static constexpr int TDW_W = DEC ? blob::DECGT_TRA_DW_W : blob::ENCGT_TRA_DW_W;
    static constexpr int TDW_B = DEC ? blob::DECGT_TRA_DW_B : blob::ENCGT_TRA_DW_B;
    static constexpr int TPW_W = DEC ? blob::DECGT_TRA_PW_W : blob::ENCGT_TRA_PW_W;
    static constexpr int TPW_B = DEC ? blob::DECGT_TRA_PW_B : blob::ENCGT_TRA_PW_B;
    static constexpr int SLOPE = DEC ? blob::DECGT_SLOPE : blob::ENCGT_SLOPE;
};

// ---- pieces shared with the streaming kernel (same operation order => bit-identical results) ----
// point conv 1: p[16] = PReLU(b + W1^T x1), x1 = channels 0..7 of (x [+ skip])
template <bool DEC, class PB>
__device__ __forceinline__ void gt_pc1(const PB& prm, const float (&x1)[8], float (&p)[16]) {
    using O = GtOff<DEC>;
#pragma unroll
    for (int c = 0; c < 16; ++c) p[c] = prm.w[O::PC1_B + c];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci)
#pragma unroll
        for (int c = 0; c < 16; c += 2) fma_pair(p[c], p[c + 1], x1[ci], prm.w[O::PC1_W + ci * 16 + c], prm.w[O::PC1_W + ci * 16 + c + 1]);
    const float a1 = prm.w[O::SLOPE + 0];
#pragma unroll
    for (int c = 0; c < 16; ++c) p[c] = prelu(p[c], a1);      // scalar on purpose: pairing here costs the GTConv kernels spills
}
// one (kf, ci-quad) step of the 3x3 conv for PS slots: in[j] = tap row j (slot s uses rows s, s+1, s+2 for kt = 0,1,2)
template <bool DEC, int PS, class PB>
__device__ __forceinline__ void gt_mid_step(const PB& prm, int kf, int qd, const float4 (&in)[PS + 2], float (&acc)[PS][16]) {
    using O = GtOff<DEC>;
    if (DEC) {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int wb = O::MID_W + ((kt * 3 + kf) * 16 + 4 * qd + i) * 16;
#pragma unroll
                for (int c = 0; c < 16; c += 2) {
                    const float w0 = prm.w[wb + c], w1 = prm.w[wb + c + 1];
#ifdef GTCRNB_DECMID_SCALAR
#pragma unroll
                    for (int s = 0; s < PS; ++s) { acc[s][c] = fmaf(f4get(in[s + kt], i), w0, acc[s][c]); acc[s][c + 1] = fmaf(f4get(in[s + kt], i), w1, acc[s][c + 1]); }
#else
#pragma unroll
                    for (int s = 0; s < PS; ++s) fma_pair(acc[s][c], acc[s][c + 1], f4get(in[s + kt], i), w0, w1);
#endif
                }
            }
        }
    } else {
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
#pragma unroll
            for (int i = 0; i < 4; i += 2) {
                const float w0 = prm.w[O::MID_W + (kt * 3 + kf) * 16 + 4 * qd + i], w1 = prm.w[O::MID_W + (kt * 3 + kf) * 16 + 4 * qd + i + 1];
#pragma unroll
                for (int s = 0; s < PS; ++s)
                    fma_pair2(acc[s][4 * qd + i], acc[s][4 * qd + i + 1], f4get(in[s + kt], i), f4get(in[s + kt], i + 1), w0, w1);
            }
        }
    }
}
// PReLU, point conv 2 (16 -> 8, BN folded): q[8]
template <bool DEC, class PB>
__device__ __forceinline__ void gt_pc2(const PB& prm, float (&d)[16], float (&q)[8]) {
    using O = GtOff<DEC>;
    const float a2 = prm.w[O::SLOPE + 1];
#pragma unroll
    for (int c = 0; c < 16; ++c) d[c] = prelu(d[c], a2);
#pragma unroll
    for (int c = 0; c < 8; ++c) q[c] = prm.w[O::PC2_B + c];
#pragma unroll
    for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int c = 0; c < 8; c += 2) fma_pair(q[c], q[c + 1], d[ci], prm.w[O::PC2_W + ci * 8 + c], prm.w[O::PC2_W + ci * 8 + c + 1]);
}
// TRALite gate of channel c from the energies of frames t-2, t-1, t (e0, e1, e2 point at 8 floats each).  Evaluated by ONE
// thread per (frame, channel) and shared through shared memory: doing it per bin-thread (33x redundant, 8 expf each) cost 22 %
// of the GTConv kernels' stall samples (ncu r1_v8).
template <bool DEC, class PB>
__device__ __forceinline__ float gt_gate1(const PB& prm, const float* e0, const float* e1, const float* e2, int c) {
    using O = GtOff<DEC>;
    float v = prm.w[O::TPW_B + c];
#pragma unroll
    for (int ci = 0; ci < 8; ++ci) {
        float y = prm.w[O::TDW_B + ci];
        y = fmaf(prm.w[O::TDW_W + 0 * 8 + ci], e0[ci], y);
        y = fmaf(prm.w[O::TDW_W + 1 * 8 + ci], e1[ci], y);
        y = fmaf(prm.w[O::TDW_W + 2 * 8 + ci], e2[ci], y);
        v = fmaf(prm.w[O::TPW_W + ci * 8 + c], y, v);
    }
    return 1.0f / (1.0f + expf(-v));
}
// fixed-order energy sum over the 33 bins of one frame slot: E2 [33][8] -> mean
__device__ __forceinline__ float gt_energy(const float* e2_slot_c) {
    float s = 0.f;
#pragma unroll
    for (int ff = 0; ff < kF2; ++ff) s += e2_slot_c[ff * 8];
    return s / 33.0f;
}

template <bool DEC, int PS, int G, bool SMEM_W>
__global__ void __launch_bounds__(gt_threads(G), gt_min_blocks(PS, G))
k_gtconv(const float* __restrict__ xin, const float* __restrict__ skip, float* __restrict__ out,
         const __grid_constant__ ParamBlock<gt_chunk<DEC>()> prm_c, int T, int seg_len, int B) {
    constexpr int NT = G * PS, R = NT + 2, GT_THREADS = gt_threads(G), GT_ACTIVE = G * kF2;
    static_assert(NT * 8 <= GT_THREADS, "phase 3 needs one thread per (slot, channel)");
    extern __shared__ float4 smem4[];
    float4* Pb = smem4;                                               // ring: R rows x GT_ROW4
    float* E2 = reinterpret_cast<float*>(Pb + R * GT_ROW4);           // [NT][33][8] q^2 of the current tile
    float* Ev = E2 + NT * kF2 * 8;                                    // ring: R rows x 8 TRA energies

    const int tid = threadIdx.x;
    // Balanced segmentation (round 2): the (utterance, frame) space is flattened and CTA c walks frames [c seg_len, (c + 1) seg_len)
    // of it -- every CTA gets the same amount of work whatever B is (with one grid row per utterance and equal segments the last
    // wave was up to 30 % empty).  A range that crosses an utterance boundary restarts with zero history at t = 0 of the next
    // utterance; a range that starts inside an utterance warms up on the 4 preceding frames as before (bit-identical).
    // (The FFMA decoder instantiation keeps one grid row per utterance and one segment per CTA: with the extra loop level ptxas fell
    //  back to per-thread-indexed constant loads and spills for its 2,304-weight contraction -- tests/test_build_static.py.)
    const int64_t fl_begin = DEC ? (int64_t)blockIdx.y * T + (int64_t)blockIdx.x * seg_len : (int64_t)blockIdx.x * seg_len;
    const int64_t fl_end = DEC ? min(fl_begin + seg_len, (int64_t)(blockIdx.y + 1) * T) : min(fl_begin + seg_len, (int64_t)B * T);
    // weight source: kernel parameter (constant bank) or a shared-memory copy read with broadcast LDS.128
    float* Gt = Ev + R * 8;                                           // [NT][8] TRALite gates of the current tile
    float* wsm = Gt + NT * 8;
    if (SMEM_W)
        for (int i = tid; i < gt_chunk<DEC>(); i += GT_THREADS) wsm[i] = prm_c.w[i];
    const SmemWeights prm_s{static_cast<const float*>(__builtin_assume_aligned(wsm, 16))};
    const auto& prm = pick_weights<SMEM_W>(prm_s, prm_c);
    const bool act = tid < GT_ACTIVE;
    const int g = tid / kF2, f = tid - g * kF2;

    int64_t fl = fl_begin;
    if (DEC && fl >= fl_end) return;
    do {
    const int b = DEC ? (int)blockIdx.y : (int)(fl / T);
    const int s0 = DEC ? (int)(blockIdx.x * seg_len) : (int)(fl - (int64_t)b * T), s1 = (int)min((int64_t)T, s0 + (fl_end - fl));
    fl += s1 - s0;
    const int tw0 = max(0, s0 - GT_WARMUP);
    __syncthreads();                                                  // the previous piece's last readers of the tile buffers are done
    zero_f4(reinterpret_cast<float*>(Pb), gt_smem_floats(PS, G), tid, GT_THREADS);
    __syncthreads();

    const int64_t ubase = (int64_t)b * T * kF2 * 16;
    int rb = 2;                                                       // ring row of slot 0 (rows 0,1: zero history)

    for (int t0 = tw0; t0 < s1; t0 += NT) {
        int rrow[PS + 2];                                             // ring rows of frames (slot0 - 2 .. slot0 + PS - 1) of this thread
#pragma unroll
        for (int j = 0; j < PS + 2; ++j) {
            int v = rb + g * PS + j + (R - 2);
            v -= (v >= R) ? R : 0;
            v -= (v >= R) ? R : 0;
            rrow[j] = v;
        }
        // ---- phase 1: point conv 1 of the new frames into the ring
        if (act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                const int t = t0 + g * PS + s;
                float p[16];
                if (t < s1) {
                    const float4* px = reinterpret_cast<const float4*>(xin + ubase + ((int64_t)t * kF2 + f) * 16);
                    float4 v0 = px[0], v1 = px[1];
                    if (DEC) {
                        const float4* ps = reinterpret_cast<const float4*>(skip + ubase + ((int64_t)t * kF2 + f) * 16);
                        const float4 u0 = ps[0], u1 = ps[1];
                        v0.x += u0.x; v0.y += u0.y; v0.z += u0.z; v0.w += u0.w;
                        v1.x += u1.x; v1.y += u1.y; v1.z += u1.z; v1.w += u1.w;
                    }
                    const float x1[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    gt_pc1<DEC>(prm, x1, p);
                } else {
#pragma unroll
                    for (int c = 0; c < 16; ++c) p[c] = 0.f;
                }
                float4* row = Pb + rrow[s + 2] * GT_ROW4 + f + 1;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) row[qd * 35] = make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]);
                // pull the same position of the NEXT tile towards L2 while this tile computes (phase-1 loads were 8-17 % of the
                // kernel's stall samples as long_scoreboard, ncu r1_v8)
                const int tn = t + NT;
                if (tn < s1) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + ubase + ((int64_t)tn * kF2 + f) * 16));
                    if (DEC) asm volatile("prefetch.global.L2 [%0];" ::"l"(skip + ubase + ((int64_t)tn * kF2 + f) * 16));
                }
            }
        }
        __syncthreads();
        // ---- phase 2: 3x3 conv over (t-2..t, f-1..f+1), PReLU, point conv 2, q^2
        // (slot groups lying wholly past the segment end skip the arithmetic: the last tile of an utterance is mostly empty)
        float q[PS][8];
        const bool live = act && (t0 + g * PS < s1);
        if (!live && act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
#pragma unroll
                for (int c = 0; c < 8; ++c) q[s][c] = 0.f;
                float4* e4 = reinterpret_cast<float4*>(E2 + ((g * PS + s) * kF2 + f) * 8);
                e4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
                e4[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (live) {
            float acc[PS][16];
#pragma unroll
            for (int s = 0; s < PS; ++s)
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[s][c] = prm.w[GtOff<DEC>::MID_B + c];
            // (kf, ci-quad) loop kept rolled: the body (3 kt x 4 ci x 16 co x PS FFMAs) is ~8 KB of code; fully unrolled
            // it was 32 KB per kf and the kernel stalled 23 % of the time on instruction fetch (ncu r1_v2: no_inst)
            if (DEC) {
#if GTCRNB_DEC_UNROLL == 0
#pragma unroll 1
                for (int kq = 0; kq < 12; ++kq) {
                    const int kf = kq >> 2, qd = kq & 3;
                    float4 in[PS + 2];
#pragma unroll
                    for (int j = 0; j < PS + 2; ++j) in[j] = Pb[rrow[j] * GT_ROW4 + qd * 35 + f + kf];
                    gt_mid_step<DEC, PS>(prm, kf, qd, in, acc);
                }
#elif GTCRNB_DEC_UNROLL == 1
#pragma unroll 1
                for (int kf = 0; kf < 3; ++kf) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        float4 in[PS + 2];
#pragma unroll
                        for (int j = 0; j < PS + 2; ++j) in[j] = Pb[rrow[j] * GT_ROW4 + qd * 35 + f + kf];
                        gt_mid_step<DEC, PS>(prm, kf, qd, in, acc);
                    }
                }
#else
                // forced full unroll (a plain `#pragma unroll` is declined at this body size): every weight index becomes an
                // immediate, so the loads are LDCU.128 through the uniform datapath instead of register-indexed LDC.64
                // through the indexed constant cache, whose ~0.5 requests/clk/SM ceiling bound the rolled loop (ncu r1_v12:
                // idc__request_cycles_active 62 %, pipe_adu 64 %, fma pipe 58 %)
                static_for<12>([&](auto kq_) {
                    constexpr int kq = decltype(kq_)::value, kf = kq >> 2, qd = kq & 3;
                    float4 in[PS + 2];
#pragma unroll
                    for (int j = 0; j < PS + 2; ++j) in[j] = Pb[rrow[j] * GT_ROW4 + qd * 35 + f + kf];
                    gt_mid_step<DEC, PS>(prm, kf, qd, in, acc);
                });
#endif
            } else {
#pragma unroll 1
                for (int kf = 0; kf < 3; ++kf) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        float4 in[PS + 2];
#pragma unroll
                        for (int j = 0; j < PS + 2; ++j) in[j] = Pb[rrow[j] * GT_ROW4 + qd * 35 + f + kf];
                        gt_mid_step<DEC, PS>(prm, kf, qd, in, acc);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                gt_pc2<DEC>(prm, acc[s], q[s]);
                float4* e4 = reinterpret_cast<float4*>(E2 + ((g * PS + s) * kF2 + f) * 8);
                e4[0] = make_float4(q[s][0] * q[s][0], q[s][1] * q[s][1], q[s][2] * q[s][2], q[s][3] * q[s][3]);
                e4[1] = make_float4(q[s][4] * q[s][4], q[s][5] * q[s][5], q[s][6] * q[s][6], q[s][7] * q[s][7]);
            }
        }
        __syncthreads();
        // ---- phase 3: TRALite energies (fixed summation order: bitwise deterministic / causal), then one gate per (slot, channel)
        int erow = 0;
        if (tid < NT * 8) {
            const int slot = tid >> 3, c = tid & 7;
            erow = rb + slot;
            erow -= (erow >= R) ? R : 0;
            Ev[erow * 8 + c] = gt_energy(E2 + slot * kF2 * 8 + c);
        }
        __syncthreads();
        if (tid < NT * 8) {
            const int slot = tid >> 3, c = tid & 7;
            int r1 = erow - 1; r1 += (r1 < 0) ? R : 0;
            int r2 = r1 - 1; r2 += (r2 < 0) ? R : 0;
            Gt[slot * 8 + c] = gt_gate1<DEC>(prm, Ev + r2 * 8, Ev + r1 * 8, Ev + erow * 8, c);
        }
        __syncthreads();
        // ---- phase 4: gate, channel shuffle [q0,x2_0,q1,x2_1,...], store
        if (act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                const int t = t0 + g * PS + s;
                if (t >= s0 && t < s1) {
                    const float4* px = reinterpret_cast<const float4*>(xin + ubase + ((int64_t)t * kF2 + f) * 16);
                    float4 v2 = px[2], v3 = px[3];                    // untouched half (re-read: L2 hit, saves 8*PS registers)
                    if (DEC) {
                        const float4* ps = reinterpret_cast<const float4*>(skip + ubase + ((int64_t)t * kF2 + f) * 16);
                        const float4 u2 = ps[2], u3 = ps[3];
                        v2.x += u2.x; v2.y += u2.y; v2.z += u2.z; v2.w += u2.w;
                        v3.x += u3.x; v3.y += u3.y; v3.z += u3.z; v3.w += u3.w;
                    }
                    const float4 g0 = reinterpret_cast<const float4*>(Gt + (g * PS + s) * 8)[0];
                    const float4 g1 = reinterpret_cast<const float4*>(Gt + (g * PS + s) * 8)[1];
                    float4* po = reinterpret_cast<float4*>(out + ubase + ((int64_t)t * kF2 + f) * 16);
                    po[0] = make_float4(q[s][0] * g0.x, v2.x, q[s][1] * g0.y, v2.y);
                    po[1] = make_float4(q[s][2] * g0.z, v2.z, q[s][3] * g0.w, v2.w);
                    po[2] = make_float4(q[s][4] * g1.x, v3.x, q[s][5] * g1.y, v3.y);
                    po[3] = make_float4(q[s][6] * g1.z, v3.z, q[s][7] * g1.w, v3.w);
                }
            }
        }
        rb += NT;
        rb -= (rb >= R) ? R : 0;
        // no barrier needed here: the next phase 1 only writes ring rows whose readers finished before the
        // barrier after phase 2, and every later writer of E2 / Ev / Gt is separated from this tile's readers by >= 2 barriers
    }
    } while (!DEC && fl < fl_end);
}

// ================================================================================================
// GTCN: 4 TCN blocks (dilation 1,2,4,8) fused.  reference: gtcrn_micro.py:256-336
// TCNs never mix frequency bins: every (utterance, bin) column is an independent causal recurrence along time.  The
// block input x stays in registers through all four TCNs; weights are a kernel parameter (constant bank).  Y holds the
// conv1 output of the tile for the TCN being evaluated; H[parity] keeps the last 2d rows of every TCN from the previous
// tile (the streaming tcn_cache, gtcrn_micro_stream.py:620-623), ping-ponged so saving the new history never disturbs the
// tap readers of the current one.  Segments that do not start at t = 0 warm up on the 30 preceding frames.
// ================================================================================================
constexpr int TC_WARMUP = 30;

// One TCN for P positions held by a thread.  Every weight is loaded once (LDCU) and used for P FFMAs; conv3 accumulates
// straight into x (bias + residual + sum) and the depthwise stage overwrites y1 in place, so the live state is 2*P*16 floats.
template <int P, class PB>
__device__ __forceinline__ void tcn_conv1(const PB& prm, int wo, const float (&x)[P][16], float (&y1)[P][16]) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float bb = prm.w[wo + blob::TCN_B1 + c];
#pragma unroll
        for (int s = 0; s < P; ++s) y1[s][c] = bb;
    }
#pragma unroll
    for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            const float w0 = prm.w[wo + blob::TCN_W1 + ci * 16 + c], w1 = prm.w[wo + blob::TCN_W1 + ci * 16 + c + 1];
#pragma unroll
            for (int s = 0; s < P; ++s) fma_pair(y1[s][c], y1[s][c + 1], x[s][ci], w0, w1);
        }
    const float a1 = prm.w[wo + blob::TCN_SLOPE + 0];
#pragma unroll
    for (int s = 0; s < P; ++s)
#pragma unroll
        for (int c = 0; c < 16; c += 2) prelu_pair(y1[s][c], y1[s][c + 1], a1);
}
// depthwise dilated temporal conv, in place: y[c] <- PReLU(b2 + w0 * tap(t-2d) + w1 * tap(t-d) + w2 * y[c])
template <class PB>
__device__ __forceinline__ void tcn_dw(const PB& prm, int wo, const float4 (&ta)[4], const float4 (&tb)[4], float (&y)[16]) {
    const float a2 = prm.w[wo + blob::TCN_SLOPE + 1];
#pragma unroll
    for (int qd = 0; qd < 4; ++qd) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {                        // scalar on purpose: pairing these per-channel taps makes ptxas fetch ALL
            const int c = 4 * qd + i;                        // of the TCN's weights with per-thread-indexed LDC (tools/sass_check.sh), 2x slower
            float v = prm.w[wo + blob::TCN_B2 + c];
            v = fmaf(prm.w[wo + blob::TCN_W2 + c], f4get(ta[qd], i), v);
            v = fmaf(prm.w[wo + blob::TCN_W2 + 16 + c], f4get(tb[qd], i), v);
            v = fmaf(prm.w[wo + blob::TCN_W2 + 32 + c], y[c], v);
            y[c] = prelu(v, a2);
        }
    }
}
template <int P, class PB>
__device__ __forceinline__ void tcn_conv3(const PB& prm, int wo, const float (&y2)[P][16], float (&x)[P][16]) {
#pragma unroll
    for (int c = 0; c < 16; c += 2) {
        const float b0 = prm.w[wo + blob::TCN_B3 + c], b1 = prm.w[wo + blob::TCN_B3 + c + 1];
#pragma unroll
        for (int s = 0; s < P; ++s) add_pair(x[s][c], x[s][c + 1], b0, b1);
    }
#pragma unroll
    for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
            const float w0 = prm.w[wo + blob::TCN_W3 + ci * 16 + c], w1 = prm.w[wo + blob::TCN_W3 + ci * 16 + c + 1];
#pragma unroll
            for (int s = 0; s < P; ++s) fma_pair(x[s][c], x[s][c + 1], y2[s][ci], w0, w1);
        }
    const float a3 = prm.w[wo + blob::TCN_SLOPE + 2];
#pragma unroll
    for (int s = 0; s < P; ++s)
#pragma unroll
        for (int c = 0; c < 16; c += 2) prelu_pair(x[s][c], x[s][c + 1], a3);
}

// GTCN walking kernel, one WARP per (utterance, bin) column: a lane owns frame slots lane, lane+32, .. (P of them) of a
// 32*P-frame tile, keeps x in registers through the four TCNs and exchanges conv1 outputs with the other lanes through the
// warp's private shared memory -- only __syncwarp, no CTA barrier, every lane busy (the CTA-tiled v2 kernel lost 21 % of its
// stall samples to barriers and 8 % of its lanes to 264/288 occupancy).
// grid (ceil(B*33 / WC_WARPS), nseg), WC_WARPS*32 threads.
#ifndef GTCRNB_WC_WARPS
#define GTCRNB_WC_WARPS 8
#endif
constexpr int WC_WARPS = GTCRNB_WC_WARPS;           // columns (warps) per CTA
constexpr int WC_MIN_BLOCKS = WC_WARPS >= 8 ? 2 : 18 / WC_WARPS;
__host__ __device__ constexpr int wc_nt(int P) { return 32 * P; }
__host__ __device__ constexpr int wc_warp4(int P) { return 4 * wc_nt(P) + 2 * 4 * 30; }       // float4 per warp: Y[q][row] + H[par][q][30]
__host__ __device__ constexpr int wc_smem_floats(int P, bool smem_w = false) { return WC_WARPS * wc_warp4(P) * 4 + (smem_w ? 4 * blob::TCN_SIZE : 0); }

template <int P, bool SMEM_W>
__global__ void __launch_bounds__(WC_WARPS * 32, (P <= 2 ? WC_MIN_BLOCKS : 1))
k_gtcn(const float* __restrict__ xin, float* __restrict__ out, const __grid_constant__ ParamBlock<4 * blob::TCN_SIZE> prm_c,
       int T, int seg_len, int64_t ncols) {
    constexpr int NT = wc_nt(P);
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* Y = smem4 + warp * wc_warp4(P);
    float4* H = Y + 4 * NT;
    // weight source: kernel parameter (constant bank, LDCU) or a shared-memory copy (broadcast LDS)
    float* wsm = reinterpret_cast<float*>(smem4 + WC_WARPS * wc_warp4(P));
    if (SMEM_W) {
        for (int i = threadIdx.x; i < 4 * blob::TCN_SIZE; i += WC_WARPS * 32) wsm[i] = prm_c.w[i];
        __syncthreads();
    }
    const SmemWeights prm_s{static_cast<const float*>(__builtin_assume_aligned(wsm, 16))};
    const auto& prm = pick_weights<SMEM_W>(prm_s, prm_c);
    const int64_t colid = (int64_t)blockIdx.x * WC_WARPS + warp;
    if (colid >= ncols) return;                                   // whole warp exits together (no CTA barrier below)
    const int64_t b = colid / kF2;
    const int f = (int)(colid - b * kF2);
    const int s0 = blockIdx.y * seg_len, s1 = min(T, s0 + seg_len);
    if (s0 >= s1) return;
    const int tw0 = max(0, s0 - TC_WARMUP);
    for (int i = lane; i < 240; i += 32) H[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    const float* xcol = xin + (b * T * kF2 + f) * 16;
    float* ocol = out + (b * T * kF2 + f) * 16;
    int par = 0;

    for (int t0 = tw0; t0 < s1; t0 += NT, par ^= 1) {
        float x[P][16];
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const int t = t0 + lane + 32 * s;
            if (t < s1) {
                const float4* px = reinterpret_cast<const float4*>(xcol + (int64_t)t * (kF2 * 16));
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 u = px[v];
                    x[s][4 * v] = u.x; x[s][4 * v + 1] = u.y; x[s][4 * v + 2] = u.z; x[s][4 * v + 3] = u.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) x[s][c] = 0.f;
            }
        }
#if GTCRNB_TCN_UNROLL
        static_for<4>([&](auto k_) {
            constexpr int k = decltype(k_)::value;
#else
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
#endif
            const int d = 1 << k, coff = (2 << k) - 2, wo = k * blob::TCN_SIZE;   // history rows of TCN k start at 2^(k+1) - 2
            static_assert(blob::TCN_SIZE % 4 == 0, "16-byte aligned weight chunks");
            const float4* Hc = H + par * 120 + coff;                               // [q] stride 30: frames t0-2d .. t0-1
            float4* Hn = H + (par ^ 1) * 120 + coff;
            float y[P][16];
            tcn_conv1<P>(prm, wo, x, y);
#pragma unroll
            for (int s = 0; s < P; ++s)
#pragma unroll
                for (int qd = 0; qd < 4; ++qd)
                    Y[qd * NT + lane + 32 * s] = make_float4(y[s][4 * qd], y[s][4 * qd + 1], y[s][4 * qd + 2], y[s][4 * qd + 3]);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < P; ++s) {
                const int sl = lane + 32 * s;
                float4 ta[4], tb[4];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    ta[qd] = sl >= 2 * d ? Y[qd * NT + sl - 2 * d] : Hc[qd * 30 + sl];
                    tb[qd] = sl >= d ? Y[qd * NT + sl - d] : Hc[qd * 30 + sl + d];
                }
                tcn_dw(prm, wo, ta, tb, y[s]);
            }
            tcn_conv3<P>(prm, wo, y, x);
            // history for the next tile <- last 2d rows of this tile (other parity, so this tile's tap readers are undisturbed)
            for (int i = lane; i < 8 * d; i += 32) {
                const int qd = i / (2 * d), j = i - qd * 2 * d;
                Hn[qd * 30 + j] = Y[qd * NT + NT - 2 * d + j];
            }
            __syncwarp();
#if GTCRNB_TCN_UNROLL
        });
#else
        }
#endif
#pragma unroll
        for (int s = 0; s < P; ++s) {
            const int t = t0 + lane + 32 * s;
            if (t >= s0 && t < s1) {
                float4* po = reinterpret_cast<float4*>(ocol + (int64_t)t * (kF2 * 16));
#pragma unroll
                for (int v = 0; v < 4; ++v) po[v] = make_float4(x[s][4 * v], x[s][4 * v + 1], x[s][4 * v + 2], x[s][4 * v + 3]);
            }
        }
    }
}

// ================================================================================================
// BACK: (de2 + en1) -> de3 ConvT(16->16,k5,s2) +PReLU -> (+en0) -> de4 ConvT(16->2,k5,s2) + tanh
//       -> ERB expand -> complex mask * spec -> enhanced spec.
// reference: gtcrn_micro.py:441-460 (de_convs[3..4]), :69-73 (erb.bs), :478-482 (Mask), :529-530
// transposed convs in gather form: out[g] = sum_{k : g+2-k even} w[k] . x[(g+2-k)/2]
// Frames are independent: (b,t) flattened, grid ceil(B*T/8), 288 threads, 8 frames per CTA.
// ================================================================================================
constexpr int BK_NF = 8;
constexpr int BK_THREADS = 288;
#ifndef GTCRNB_AHEAD_CTAS
#define GTCRNB_AHEAD_CTAS (3 * 148)
#endif
constexpr int kAheadCtas = GTCRNB_AHEAD_CTAS;              // resident CTAs of the per-frame kernels on a B200 (3 per SM): look-ahead distance of their L2 prefetches
constexpr int BK_X4 = 4 * 35;                    // float4 per slot: [q][f + 1]
constexpr int BK_Y4 = 4 * 67;                    // float4 per slot: [q][g + 1]
constexpr int BK_M = 130 * 2;                    // floats per slot: [h][2] -- overlays X (dead after de3)
static_assert(BK_M <= BK_X4 * 4, "mask buffer must fit in the dead input tile");
constexpr int BK_FFT_FLOATS = FR_FFT_FLOATS + 256;        // FftTables + 1/envelope
constexpr int BK_W_FLOATS = blob::BACK_SIZE - blob::BACK_DE4_W;   // only de4 / inverse-ERB tables go to shared memory (non-uniform indices);
                                                                  // the de3 weights are a kernel parameter
static_assert(BK_W_FLOATS == blobf::BACK_SIZE - blobf::BACK_DE4_W, "both variants copy the same number of table floats");
__host__ __device__ constexpr int bk_smem_floats(bool to_wav, bool tcb = false) {
    return BK_W_FLOATS + (tcb ? TCW_DE3_WORDS : 0) + BK_NF * (BK_X4 * 4 + BK_Y4 * 4) + (to_wav ? BK_FFT_FLOATS : 0);
}
constexpr int TCB_PLANE = TCE_ROWS * 16;        // bytes per activation plane in k_back: [hl * 2 + ci chunk][row = slot * 35 + bin + 1][8 halves]
static_assert(4 * TCB_PLANE == BK_NF * BK_X4 * 16, "the four fp16 planes occupy exactly the fp32 X region");

// de3 for output bins of one parity (compile-time, so tap set and weight indices are warp-uniform): thread = 2 slots x 16 co.
// transposed conv in gather form: even g = 2m uses taps k = 0,2,4 on input bins m+1, m, m-1; odd g = 2m+1 taps k = 1,3 on m+1, m.
// VAR 1 (groups = 2): input channel quad qd feeds only the 8 outputs of its group (qd / 2), weights [k][ci][co within the group]
template <int ODD, int VAR, class PB>
__device__ __forceinline__ void back_de3(const PB& prm, const float4* X, float4* Y3, int sp2, int m, int64_t fi0, int64_t nframes,
                                         const float* __restrict__ en0) {
    using LY = BackL<VAR>;
    const int g = 2 * m + ODD;
    float acc0[16], acc1[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc0[c] = acc1[c] = prm.w[LY::DE3_B + c];
    const float4* x0 = X + (2 * sp2) * BK_X4;
    const float4* x1 = x0 + BK_X4;
#pragma unroll 1
    for (int j = 0; j < 3 - ODD; ++j) {
        const int colx = m + 2 - j;                            // input bin f = m + 1 - j, stored at col f + 1
        const int wb = LY::DE3_W + (2 * j + ODD) * (VAR == 1 ? 128 : 256);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float4 i0 = x0[qd * 35 + colx], i1 = x1[qd * 35 + colx];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if constexpr (VAR == 1) {
#pragma unroll
                    for (int c = 0; c < 8; c += 2) {
                        const float w0 = prm.w[wb + (4 * qd + i) * 8 + c], w1 = prm.w[wb + (4 * qd + i) * 8 + c + 1];
                        fma_pair(acc0[8 * (qd >> 1) + c], acc0[8 * (qd >> 1) + c + 1], f4get(i0, i), w0, w1);
                        fma_pair(acc1[8 * (qd >> 1) + c], acc1[8 * (qd >> 1) + c + 1], f4get(i1, i), w0, w1);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        const float w0 = prm.w[wb + (4 * qd + i) * 16 + c], w1 = prm.w[wb + (4 * qd + i) * 16 + c + 1];
                        fma_pair(acc0[c], acc0[c + 1], f4get(i0, i), w0, w1);
                        fma_pair(acc1[c], acc1[c + 1], f4get(i1, i), w0, w1);
                    }
                }
            }
        }
    }
    const float a = prm.w[LY::SLOPE + 0];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int slot = 2 * sp2 + s;
        float* acc = s ? acc1 : acc0;
        const bool ok = fi0 + slot < nframes;
        const float4* pe = reinterpret_cast<const float4*>(en0 + ((fi0 + (ok ? slot : 0)) * kF1 + g) * 16);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float4 e = pe[qd];
            Y3[slot * BK_Y4 + qd * 67 + g + 1] = make_float4(prelu(acc[4 * qd], a) + e.x, prelu(acc[4 * qd + 1], a) + e.y,
                                                             prelu(acc[4 * qd + 2], a) + e.z, prelu(acc[4 * qd + 3], a) + e.w);
        }
    }
}

// TO_WAV = false: the enhanced spectrum goes to `outp` (element strides ob, of, ot, oc).
// TO_WAV = true : the iSTFT is fused in: the masked spectrum stays in shared memory, is inverted (256-point FFT through the
//                 conjugation trick), windowed and overlap-added straight into `wav_out` [B][256 (T-1)] with vector RED.ADD.
//                 Every output sample receives exactly two addends on top of a zeroed buffer, so the result does not depend
//                 on their order (bitwise deterministic).  The caller zeroes wav_out first.
// de4 (16 -> 2) + tanh for output bins of one parity: thread = 4 slots x 2 co at bin h = 2m + ODD; every weight pair (one LDCU.64)
// feeds 8 FFMAs.  Same gather form as de3.  Result: mask on the compressed axis, M[slot][h][2].
template <int ODD, int VAR, class PB>
__device__ __forceinline__ void back_de4(const PB& prm, const float4* Y3, float* M, int m_stride, int sg, int m) {
    using LY = BackL<VAR>;
    const int h = 2 * m + ODD;
    float a0[4], a1[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) { a0[s] = prm.w[LY::DE4_B + 0]; a1[s] = prm.w[LY::DE4_B + 1]; }
    const float4* y = Y3 + (4 * sg) * BK_Y4;
#pragma unroll
    for (int j = 0; j < 3 - ODD; ++j) {
        const int colx = m + 2 - j, wb = LY::DE4_W + (2 * j + ODD) * 32;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            float4 in[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) in[s] = y[s * BK_Y4 + qd * 67 + colx];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float w0 = prm.w[wb + (4 * qd + i) * 2], w1 = prm.w[wb + (4 * qd + i) * 2 + 1];
#pragma unroll
                for (int s = 0; s < 4; ++s) fma_pair(a0[s], a1[s], f4get(in[s], i), w0, w1);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        *reinterpret_cast<float2*>(M + (4 * sg + s) * m_stride + 2 * h) = make_float2(tanhf(a0[s]), tanhf(a1[s]));
}

template <bool TO_WAV, int VAR, bool TCB = false>
__global__ void __launch_bounds__(BK_THREADS, 3) k_back(const float* __restrict__ de2, const float* __restrict__ en1,
                                                        const float* __restrict__ en0, const float* __restrict__ spec,
                                                        int64_t sb, int64_t sf, int64_t st, int64_t sc,
                                                        float* __restrict__ outp, int64_t ob, int64_t of, int64_t ot, int64_t oc,
                                                        float* __restrict__ wav_out, int64_t wav_stride,
                                                        int T, int64_t nframes, const float* __restrict__ blobp,
                                                        const __grid_constant__ ParamBlock<BackL<VAR>::SIZE> prm, const uint32_t* __restrict__ tcw = nullptr) {
    static_assert(!TCB || VAR == 0, "tensor-core de3: GTCRN-Micro");
    using LY = BackL<VAR>;
    constexpr int BK_W_OFF = LY::DE4_W;
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) uint64_t tc_bars[3];
    __shared__ uint32_t tmem_base_s;
    float* Wsm = reinterpret_cast<float*>(smem4);
    float* W = Wsm - BK_W_OFF;                       // so that W[LY::xxx] addresses the copied sub-range
    float* Btc = Wsm + BK_W_FLOATS;                  // TCB: de3 weight matrices (k_prep_ends_tc layout)
    float4* X = reinterpret_cast<float4*>(Btc + (TCB ? TCW_DE3_WORDS : 0));
    float4* const X4all = X;
    float4* Y3 = X + BK_NF * BK_X4;
    float* M = reinterpret_cast<float*>(X);          // stride BK_X4 * 4 floats per slot
    constexpr int M_STRIDE = BK_X4 * 4;
    const int tid = threadIdx.x;
    const int64_t fi0 = (int64_t)blockIdx.x * BK_NF;
    FBT(1, 0);

    // ---- a: x = de2 + en1 into X.  The two input tiles of a CTA are contiguous 16.9 KB blocks and the tables are the same for
    //      every CTA: ONE thread moves all of it with 1-D bulk copies (TMA, cp.async.bulk -> SASS UBLKCP) into shared memory -- the
    //      tiles into the Y3 region, which is dead until de3 writes it -- and everybody waits on the copies' mbarrier.  [Per-thread
    //      LDG.128 + cp.async of the same bytes: 2,700 requests per CTA through the LSU; the clock64 trace (tools/gpu_fb_trace.sh)
    //      showed this phase taking 9,800 of a CTA's 34,000 cycles.]  The operands this CTA needs much later (en0 skip rows for de3,
    //      the spectrum for the mask) are pulled towards L2 now, and so are the tiles of the CTA one residency wave ahead.
    __shared__ __align__(8) uint64_t ld_bar;
    const int nf_a = (int)min((int64_t)BK_NF, nframes - fi0);                        // frames of this CTA (contiguous rows)
    if (tid == 0) {
        tc::mbar_init(&ld_bar, 1);
        tc::fence_barrier_init();
        const uint32_t tile_bytes = (uint32_t)nf_a * kF2 * 64, fft_bytes = TO_WAV ? 2048u + 2064u + 2048u + 1024u : 0u;
        tc::mbar_expect_tx(&ld_bar, 2 * tile_bytes + BK_W_FLOATS * 4 + fft_bytes + (TCB ? TCW_DE3_WORDS * 4 : 0));
        if (TCB) tc::bulk_g2s(Btc, tcw + TCW_EN1_WORDS, TCW_DE3_WORDS * 4, &ld_bar);
        tc::bulk_g2s(Y3, de2 + fi0 * kF2 * 16, tile_bytes, &ld_bar);
        tc::bulk_g2s(reinterpret_cast<char*>(Y3) + BK_NF * kF2 * 64, en1 + fi0 * kF2 * 16, tile_bytes, &ld_bar);
        tc::bulk_g2s(Wsm, blobp + LY::OFF + BK_W_OFF, BK_W_FLOATS * 4, &ld_bar);     // de4 / inverse-ERB tables
        if (TO_WAV) {                                                                // iSTFT tables: FftTables order + 1 / envelope
            FftTables& tb0 = *reinterpret_cast<FftTables*>(Wsm + bk_smem_floats(false, TCB));
            const float* fb = blobp + blob::FFT_OFF;
            tc::bulk_g2s(tb0.tw256, fb + blob::FFT_TW256, 2048, &ld_bar);
            tc::bulk_g2s(tb0.tw512, fb + blob::FFT_TW512, 2064, &ld_bar);
            tc::bulk_g2s(tb0.win, fb + blob::FFT_WIN, 2048, &ld_bar);
            tc::bulk_g2s(reinterpret_cast<float*>(&tb0) + FR_FFT_FLOATS, fb + blob::FFT_INV_ENV, 1024, &ld_bar);
        }
    }
    static_assert(BK_NF * BK_Y4 * 16 >= 2 * BK_NF * kF2 * 64 && sizeof(FftTables) == 2048 + 2064 + 2048, "staging fits the Y3 region; table sizes");
    if (TCB) {
        if (tid >= 32 && tid < 64) { tc::tmem_alloc(&tmem_base_s, 128); tc::fence_before(); }     // 3 tiles x 32 accumulator columns (even | odd bins)
        if (tid == 64) { for (int i = 0; i < 3; ++i) tc::mbar_init(&tc_bars[i], 1); tc::fence_async_smem(); }
    }
    const bool loader = tid < (BK_NF / 2) * kF2;
    const int sp2_a = tid / kF2, f_a = tid - sp2_a * kF2;
#ifndef GTCRNB_NO_AHEAD_PREFETCH
    {
        // CTAs start in blockIdx order as slots free up, so CTA blockIdx + kAheadCtas begins roughly when this one ends: pull its
        // tiles from HBM into L2 now
        const int64_t fa = fi0 + (int64_t)kAheadCtas * BK_NF;
        if (fa < nframes && tid * 128 < BK_NF * kF2 * 64) {
            const int64_t left = min((int64_t)BK_NF, nframes - fa) * kF2 * 64;
            if (tid * 128 < left) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(de2 + fa * kF2 * 16) + tid * 128));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(en1 + fa * kF2 * 16) + tid * 128));
            }
        }
        // (measured: look-ahead 222 or 444 CTAs 3.25 ms, 888 CTAs 3.38; also pulling the ahead CTA's en0 / spectrum rows 3.43)
    }
#endif
    {
        static_assert(BK_THREADS * 128 >= BK_NF * kF1 * 64 && BK_THREADS * 128 >= BK_NF * kSpecFrame * 4, "one 128-byte line per thread covers the rows");
        if (tid * 128 < nf_a * kF1 * 64)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(en0 + fi0 * kF1 * 16) + tid * 128));
        if (TO_WAV && tid * 128 < nf_a * kSpecFrame * 4)                             // fused mode: spec is the contiguous [frame][257][2] buffer
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(spec + fi0 * kSpecFrame) + tid * 128));
    }
    for (int i = tid; i < BK_NF * 8; i += BK_THREADS) {            // zero pad columns 0 and 34 of X (those of Y3 follow once the staging is consumed)
        const int slot = i >> 3, q = (i >> 1) & 3, side = i & 1;
        if (TCB) reinterpret_cast<uint4*>(X)[q * TCE_ROWS + slot * 35 + (side ? 34 : 0)] = make_uint4(0u, 0u, 0u, 0u);   // q = plane (hl * 2 + chunk)
        else X[slot * BK_X4 + q * 35 + (side ? 34 : 0)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();                                               // the mbarrier is initialised for everybody
    tc::mbar_wait(&ld_bar, 0);
    if (loader) {
        const float4* sa = reinterpret_cast<const float4*>(Y3);
        const float4* sb_ = sa + BK_NF * kF2 * 4;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int slot = 2 * sp2_a + s;
            const int o = (slot * kF2 + f_a) * 4;
            if (TCB) {
                // A operand of the de3 MMAs: x = de2 + en1 as fp16 hi' / lo rows (row = slot * 35 + bin + 1) of four planes
                float xs[16];
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    const float4 u = sa[o + qd], v = sb_[o + qd];
                    xs[4 * qd] = u.x + v.x; xs[4 * qd + 1] = u.y + v.y; xs[4 * qd + 2] = u.z + v.z; xs[4 * qd + 3] = u.w + v.w;
                }
                uint4 hi[2], lo[2];
                split_f16_row(xs, hi, lo);
                uint4* pl = reinterpret_cast<uint4*>(X) + slot * 35 + f_a + 1;
                pl[0] = hi[0]; pl[TCE_ROWS] = hi[1]; pl[2 * TCE_ROWS] = lo[0]; pl[3 * TCE_ROWS] = lo[1];
            } else {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float4 u = sa[o + qd], v = sb_[o + qd];
                X[slot * BK_X4 + qd * 35 + f_a + 1] = make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w);
            }
            }
        }
    }
    if (TCB) tc::fence_async_smem();                               // the planes are read by the tensor core (async proxy)
    FBT(1, 1);
    __syncthreads();
    FBT(1, 2);
    // ---- b: de3: even output bins on warps 0-4 (132 of 160 lanes), odd bins on warps 5-8 (128 lanes)
    for (int i = tid; i < BK_NF * 8; i += BK_THREADS) {            // pad columns 0 and 66 of Y3 (the staged tiles are consumed)
        const int slot = i >> 3, q = (i >> 1) & 3, side = i & 1;
        Y3[slot * BK_Y4 + q * 67 + (side ? 66 : 0)] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if constexpr (TCB) {
        // transposed conv in gather form on the tensor cores: tap pair j reads the planes shifted by 2 - j rows; N = 32 accumulator
        // columns = [even bins g = 2m | odd bins g = 2m + 1] x 16 channels; three MMAs per tap pair: hi' W_hi, hi' W_lo, lo W_hi
        tc::fence_after();
        const uint32_t tm = tmem_base_s;
        if (tid == 0) {
            constexpr uint32_t id32 = tc::make_idesc_f16(128, 32);
            const uint32_t a_base = tc::smem_u32(X), b_base = tc::smem_u32(Btc);
#pragma unroll 1
            for (int tile = 0; tile < 3; ++tile) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const uint32_t a_hi = a_base + (uint32_t)((128 * tile + 2 - j) * 16);
                    const uint64_t da_hi = tc::make_desc(a_hi, TCB_PLANE, 128), da_lo = tc::make_desc(a_hi + 2 * TCB_PLANE, TCB_PLANE, 128);
                    const uint32_t bj = b_base + j * 3072;
                    tc::mma_f16_ss(tm + 32 * tile, da_hi, tc::make_desc(bj, 512, 128), id32, j > 0);
                    tc::mma_f16_ss(tm + 32 * tile, da_hi, tc::make_desc(bj + 1024, 512, 128), id32, 1);
                    tc::mma_f16_ss(tm + 32 * tile, da_lo, tc::make_desc(bj + 2048, 512, 128), id32, 1);
                }
                tc::commit(&tc_bars[tile]);
            }
        }
        const int warp = tid >> 5, lane = tid & 31;
        {
            const int tile = warp < 8 ? warp >> 2 : 2, lg = warp & 3;
            tc::mbar_wait(&tc_bars[tile], 0);
            tc::fence_after();
            const int r = 128 * tile + 32 * lg + lane, slot = r / 35, m = r - slot * 35;
            const bool ok = r < TCE_ROWS && m < kF2 && fi0 + slot < nframes;
            const uint32_t ta = tm + ((uint32_t)(lg * 32) << 16) + 32 * tile;
            const float winv = Btc[TCW_DE3_WORDS - 4], a = prm.w[LY::SLOPE + 0];
#pragma unroll
            for (int odd = 0; odd < 2; ++odd) {
                float d[16];
                tc::tmem_ld16_nowait(ta + 16 * odd, d);
                tc::tmem_wait_ld(d);
                const int g = 2 * m + odd;
                if (ok && g < kF1) {
                    const float4* pe = reinterpret_cast<const float4*>(en0 + ((fi0 + slot) * kF1 + g) * 16);
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float4 e = pe[qd];
                        float v[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = prelu(fmaf(d[4 * qd + i], winv, prm.w[LY::DE3_B + 4 * qd + i]), a);
                        Y3[slot * BK_Y4 + qd * 67 + g + 1] = make_float4(v[0] + e.x, v[1] + e.y, v[2] + e.z, v[3] + e.w);
                    }
                }
            }
        }
        tc::fence_before();
    } else {
    if (tid < 160) {
        if (tid < (BK_NF / 2) * 33) back_de3<0, VAR>(prm, X, Y3, tid / 33, tid % 33, fi0, nframes, en0);
    } else {
        const int r = tid - 160;
        back_de3<1, VAR>(prm, X, Y3, r >> 5, r & 31, fi0, nframes, en0);
    }
    }
    FBT(1, 3);
    __syncthreads();
    if (TCB && tid >= 32 && tid < 64) tc::tmem_dealloc(tmem_base_s, 128);     // every accumulator row has been read
    FBT(1, 4);
    // fused iSTFT: warp w will mask and invert frame slot w; its 257 spectrum bins are requested NOW, so that their L2 latency runs
    // under the de4 phase instead of opening the mask phase (clock64 trace: mask phase 4,600-7,000 of a CTA's 32,000 cycles)
    float2 xsp[9];
    if (TO_WAV) {
        const int warp_ = tid >> 5, lane_ = tid & 31;
        const int64_t fi_ = fi0 + warp_;
        const bool valid_ = warp_ < BK_NF && fi_ < nframes;
        const float2* sp_ = reinterpret_cast<const float2*>(spec + (valid_ ? fi_ : fi0) * kSpecFrame);   // fused mode: contiguous [frame][257][2]
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int i = lane_ + 32 * q;
            xsp[q] = (valid_ && i <= 256) ? sp_[i] : make_float2(0.f, 0.f);
        }
    }
    // ---- c: de4 (16 -> 2) + tanh -> mask on the 129 compressed bins: even bins (65) on warps 0-4, odd bins (64) on warps 5-8
    static_assert(BK_NF == 8, "de4 mapping assumes two groups of four slots");
    if (tid < 160) {
        if (tid < 130) back_de4<0, VAR>(prm, Y3, M, M_STRIDE, tid / 65, tid % 65);
    } else {
        const int r = tid - 160;
        back_de4<1, VAR>(prm, Y3, M, M_STRIDE, r >> 6, r & 63);
    }
    FBT(1, 5);
    __syncthreads();
    FBT(1, 6);
    // ---- d: ERB expand + complex ratio mask (+ e: fused iSTFT).  mask(i) on the 257 spectrum bins from the 129 compressed ones.
    const int* ib0 = reinterpret_cast<const int*>(W + LY::IERB_B0);
    const float* iw = W + LY::IERB_W;
    auto mask_at = [&](const float* ms, int i, float& m0, float& m1) {
        if (i < 65) {
            m0 = ms[2 * i];
            m1 = ms[2 * i + 1];
        } else {
            const int j = i - 65, bb = 65 + ib0[j];
            const float w0 = iw[2 * j], w1 = iw[2 * j + 1];
            m0 = w0 * ms[2 * bb] + w1 * ms[2 * bb + 2];
            m1 = w0 * ms[2 * bb + 1] + w1 * ms[2 * bb + 3];
        }
    };
    if (!TO_WAV) {
        for (int idx = tid; idx < BK_NF * kBins; idx += BK_THREADS) {
            const int slot = idx / kBins, i = idx - slot * kBins;
            const int64_t fi = fi0 + slot;
            if (fi >= nframes) continue;
            const int64_t b = fi / T, t = fi - b * T;
            float m0, m1;
            mask_at(M + slot * M_STRIDE, i, m0, m1);
            const float* p = spec + b * sb + t * st + (int64_t)i * sf;
            const float re = p[0], im = p[sc];
            float* o = outp + b * ob + t * ot + (int64_t)i * of;
            o[0] = re * m0 - im * m1;
            o[oc] = im * m0 + re * m1;
        }
    } else {
        // warp w owns frame slot w from here on: mask its spectrum into shared memory, invert it, window, overlap-add.
        // Per slot the dead Y3 region (1072 floats) holds [0,514) spectrum | [544,1056) FFT buffer A; FFT buffer B (512 floats)
        // lives in the slot's X region -- but that region still holds the mask M until this warp has consumed it, so B is only
        // touched after the __syncwarp that follows the masking loop.
        static_assert(BK_Y4 * 4 >= 544 + 512 && BK_X4 * 4 >= 512, "fused iSTFT buffers must fit in the dead slot regions");
        FftTables& tb = *reinterpret_cast<FftTables*>(Wsm + bk_smem_floats(false, TCB));
        const float* inv_env = reinterpret_cast<const float*>(&tb) + FR_FFT_FLOATS;
        const int warp = tid >> 5, lane = tid & 31;
        if (warp < BK_NF) {
            const int slot = warp;
            const int64_t fi = fi0 + slot;
            const bool valid = fi < nframes;
            int64_t b = 0; int t = 0;
            if (valid) { b = fi / T; t = (int)(fi - b * T); }
            float* base = reinterpret_cast<float*>(Y3 + slot * BK_Y4);
            float2* Ys = reinterpret_cast<float2*>(base);       // masked spectrum of this slot
            float2* A = reinterpret_cast<float2*>(base + 544);
            float2* Bf = reinterpret_cast<float2*>(reinterpret_cast<float*>(X4all + slot * BK_X4));
            {
                const float2 (&x)[9] = xsp;
                const float* ms = M + slot * M_STRIDE;
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                    const int i = lane + 32 * q;
                    if (i <= 256) {
                        float m0, m1;
                        mask_at(ms, i, m0, m1);
                        Ys[i] = make_float2(x[q].x * m0 - x[q].y * m1, x[q].y * m0 + x[q].x * m1);
                    }
                }
            }
            __syncwarp();
            FBT(1, 7);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int k = lane + 32 * q;                        // 0..255
                float2 xk = Ys[k], xc = Ys[256 - k];
                xc.y = -xc.y;
                if (k == 0) { xk.y = 0.f; xc.y = 0.f; }             // imag of DC / Nyquist ignored (C2R semantics)
                const float2 w = tb.tw512[k];
                const float dx = xk.x - xc.x, dy = xk.y - xc.y;
                const float a = w.x * dx + w.y * dy, bq = w.x * dy - w.y * dx;   // conj(w) * D
                A[k] = valid ? make_float2(0.5f * (xk.x + xc.x) - 0.5f * bq, -(0.5f * (xk.y + xc.y) + 0.5f * a)) : make_float2(0.f, 0.f);
            }
            __syncwarp();
            FBT(1, 8);
            fft256_warp(A, Bf, tb.tw256, lane);
            FBT(1, 9);
            if (valid) {
                float* o = wav_out + b * wav_stride + (int64_t)(t - 1) * kHop;      // frame t covers output samples 256 (t-1) .. 256 (t+1)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int m = lane + 32 * q;                    // frame samples 2m, 2m+1
                    const float2 z = A[m];
                    const int i = 2 * m;
                    const bool keep = i < 256 ? t > 0 : t < T - 1;  // first half of frame 0 / second half of the last frame are trimmed
                    if (keep) {
                        const float v0 = tb.win[i] * (z.x * (1.0f / 256.0f)) * inv_env[i & 255];
                        const float v1 = tb.win[i + 1] * (-z.y * (1.0f / 256.0f)) * inv_env[(i + 1) & 255];
                        atomicAdd(reinterpret_cast<float2*>(o + i), make_float2(v0, v1));
                    }
                }
            }
        }
    }
    FBT(1, 10);
}

// ================================================================================================
// PCM16 <-> float32 on the device (wav files are PCM16: the host pipeline can ship 2 bytes per sample over PCIe).
// Decode x / 32768 (what soundfile / the reference's sf.read(dtype="float32") produce, infer.py:52); encode floor(y * 32768)
// clamped, the rule the reference's shipped *_enh.wav files follow (tests/golden/pairs.npz: 99.98 % exact, rint 50 %).
// ================================================================================================
__global__ void __launch_bounds__(256) k_pcm16_to_f32(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = ((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? n / 8 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const int4 v = reinterpret_cast<const int4*>(in)[i];
        const int w[4] = {v.x, v.y, v.z, v.w};
        float o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[2 * k] = (float)(short)(w[k] & 0xffff) * (1.0f / 32768.0f);
            o[2 * k + 1] = (float)(short)(w[k] >> 16) * (1.0f / 32768.0f);
        }
        reinterpret_cast<float4*>(out)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (float)in[i] * (1.0f / 32768.0f);
}
__device__ __forceinline__ int pcm16_of(float v) { return (int)fminf(fmaxf(floorf(v * 32768.0f), -32768.0f), 32767.0f); }
__global__ void __launch_bounds__(256) k_f32_to_pcm16(const float* __restrict__ in, int16_t* __restrict__ out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = ((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? n / 8 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        const float4 a = reinterpret_cast<const float4*>(in)[2 * i], b = reinterpret_cast<const float4*>(in)[2 * i + 1];
        int4 v;
        v.x = (pcm16_of(a.x) & 0xffff) | (pcm16_of(a.y) << 16);
        v.y = (pcm16_of(a.z) & 0xffff) | (pcm16_of(a.w) << 16);
        v.z = (pcm16_of(b.x) & 0xffff) | (pcm16_of(b.y) << 16);
        v.w = (pcm16_of(b.z) & 0xffff) | (pcm16_of(b.w) << 16);
        reinterpret_cast<int4*>(out)[i] = v;
    }
    for (int64_t i = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int16_t)pcm16_of(in[i]);
}

// ================================================================================================
// On-device quality metrics of (reference, enhanced) waveform pairs: SI-SNR and SDR exactly as the reference's evaluation
// script defines them (eval/eval_intrusive_metrics.py:74-82 sisnr_metric, :85-94 sdr_metric), from five raw moments per
// utterance accumulated in fp64 with a fixed reduction order (deterministic).  grid B, 256 threads; out [B][2] = (si_snr, sdr) dB.
// ================================================================================================
__global__ void __launch_bounds__(256) k_pair_metrics(const float* __restrict__ ref, int64_t ref_stride, const float* __restrict__ inf,
                                                      int64_t inf_stride, int64_t L, float* __restrict__ out) {
    __shared__ double red[8][5];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* r = ref + (int64_t)blockIdx.x * ref_stride;
    const float* x = inf + (int64_t)blockIdx.x * inf_stride;
    double s[5] = {0, 0, 0, 0, 0};                       // sum r, sum x, sum r^2, sum x^2, sum r x
    for (int64_t i = tid; i < L; i += 256) {
        const double a = r[i], b = x[i];
        s[0] += a; s[1] += b; s[2] += a * a; s[3] += b * b; s[4] += a * b;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_down_sync(0xffffffffu, s[k], o);
        if (lane == 0) red[warp][k] = s[k];
    }
    __syncthreads();
    if (tid == 0) {
        double t[5];
        for (int k = 0; k < 5; ++k) { t[k] = 0; for (int w = 0; w < 8; ++w) t[k] += red[w][k]; }
        const double n = (double)L, mr = t[0] / n, mx = t[1] / n;
        const double rr = t[2] - n * mr * mr, xx = t[3] - n * mx * mx, rx = t[4] - n * mr * mx;   // centred second moments
        const double a = rx / (rr + n * 1e-8);            // np.sum(ref**2 + 1e-8)
        const double tgt = a * a * rr, res = xx - 2.0 * a * rx + a * a * rr;
        out[2 * blockIdx.x] = (float)(10.0 * log10((tgt + 1e-8) / (fmax(res, 0.0) + 1e-8)));
        const double err = xx - 2.0 * rx + rr;            // sum (inf_c - ref_c)^2
        out[2 * blockIdx.x + 1] = (float)(10.0 * log10((rr + 1e-8) / (fmax(err, 0.0) + 1e-8)));
    }
}

}  // namespace gk
