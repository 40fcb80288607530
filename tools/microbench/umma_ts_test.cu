// tcgen05 TF32 with the A operand in TENSOR MEMORY (TS form): correctness, instruction cost vs N, and the full
// registers -> TMEM -> MMA -> TMEM -> registers round trip that the GTCN / decoder-GTConv tensor-core kernels are built on.
//
// Why: round 1 measured the SS form (A in shared memory, no swizzle) at ~55 cycles per M128 x N<=64 x K8 MMA regardless of N
// (profiles/umma_tf32_test_r1.txt) -- the shared-memory read of A dominates, which made N = 16 contractions no faster than FFMA.
// With thread = position (row) the activations are already in registers, one row per thread, which is exactly the TMEM layout of
// an M = 128 A operand (row m in lane m, element k in column k): tcgen05.st writes A without touching shared memory at all.
//
//   test 1  TS correctness (K = 16 / N = 16 and K = 48 / N = 48), exact tf32 inputs, plain fp32 inputs (does the hardware truncate
//           or round the low 13 bits?), 3xTF32 split vs fp64 truth
//   test 2  back-to-back MMA cost, SS vs TS, N = 16 .. 256
//   test 3  full-chip round trips: split + tcgen05.st, fence, 128-thread barrier, 3*K/8 MMAs, commit, mbarrier wait, tcgen05.ld;
//           NWG warpgroups per CTA x CTAs per SM, with and without the MMAs (st / ld / sync cost alone)
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_ts_test umma_ts_test.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "../../gtcrn_micro_b200/csrc/tc_util.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__host__ __device__ constexpr uint32_t pow2_cols(uint32_t n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : n <= 256 ? 256 : 512; }

// ---------------------------------------------------------------------------------------------------------------------------
// test 1: D[128][N] = A[128][K] * B[K][N], A through tcgen05.st, B as no-swizzle K-major planes.  mode 0: one product, operands as
// given; mode 1: 3xTF32 split
// ---------------------------------------------------------------------------------------------------------------------------
template <int K, int N>
__global__ void __launch_bounds__(128) k_ts_correct(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, int mode) {
    __shared__ __align__(128) float Bhi[K / 4][N][4];
    __shared__ __align__(128) float Blo[K / 4][N][4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    constexpr uint32_t A_HI = 0, A_LO = K, D0 = 2 * K, COLS = pow2_cols(2 * K + N);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < K * N; i += 128) {
        const int k = i / N, n = i % N;
        const float v = B[i];
        const float hi = mode ? tc::tf32_hi(v) : v;
        Bhi[k / 4][n][k % 4] = hi;
        Blo[k / 4][n][k % 4] = v - hi;
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base, COLS);
    if (tid == 0) tc::mbar_init(&bar, 1);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tb = tmem_base, lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int c = 0; c < K / 16; ++c) {
        float hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float v = A[tid * K + 16 * c + j];
            hi[j] = mode ? tc::tf32_hi(v) : v;
            lo[j] = v - hi[j];
        }
        tc::tmem_st16(tb + lane_base + A_HI + 16 * c, hi);
        tc::tmem_st16(tb + lane_base + A_LO + 16 * c, lo);
    }
    tc::tmem_wait_st();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after();
        const uint32_t idesc = tc::make_idesc_tf32(128, N);
        uint32_t acc = 0;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t bh = tc::make_desc(tc::smem_u32(&Bhi[2 * ks][0][0]), N * 16, 128);
            const uint64_t bl = tc::make_desc(tc::smem_u32(&Blo[2 * ks][0][0]), N * 16, 128);
            if (mode) {
                tc::mma_tf32_ts(tb + D0, tb + A_LO + 8 * ks, bh, idesc, acc); acc = 1;     // small terms first
                tc::mma_tf32_ts(tb + D0, tb + A_HI + 8 * ks, bl, idesc, 1);
            }
            tc::mma_tf32_ts(tb + D0, tb + A_HI + 8 * ks, bh, idesc, acc); acc = 1;
        }
        tc::commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after();
#pragma unroll
    for (int c = 0; c < N / 16; ++c) {
        float v[16];
        tc::tmem_ld16_nowait(tb + lane_base + D0 + 16 * c, v);
        tc::tmem_wait_ld(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) out[tid * N + 16 * c + j] = v[j];
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, COLS);
}

// ---------------------------------------------------------------------------------------------------------------------------
// test 2: cost of nmma back-to-back M128 x N x K8 MMAs + commit + wait, A from shared memory (SS) or tensor memory (TS)
// ---------------------------------------------------------------------------------------------------------------------------
template <int N, bool TS>
__global__ void __launch_bounds__(128) k_time(int nmma, int reps, long long* cycles, float* sink) {
    extern __shared__ __align__(128) float dsm[];
    float* A = dsm;                                   // [4][160][4]
    float* B = dsm + 4 * 160 * 4;                     // [4][N][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * 160 * 4; i += 128) A[i] = 1.0f;
    for (int i = tid; i < 4 * N * 4; i += 128) B[i] = 0.5f;
    if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
    if (tid == 0) tc::mbar_init(&bar, 1);
    tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t tb = tmem_base, lane_base = (uint32_t)(warp * 32) << 16;
    {
        float one[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) one[j] = 1.0f;
        tc::tmem_st16(tb + lane_base + 0, one);
        tc::tmem_st16(tb + lane_base + 16, one);
        tc::tmem_wait_st();
    }
    tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t idesc = tc::make_idesc_tf32(128, N), td = tb + 256;
    const uint64_t da0 = tc::make_desc(tc::smem_u32(A), 160 * 16, 128);
    const uint64_t db = tc::make_desc(tc::smem_u32(B), N * 16, 128);
    uint32_t parity = 0; long long best = 1ll << 60;
    for (int rep = 0; rep < reps; ++rep) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int i = 0; i < nmma; ++i) {
                if (TS) tc::mma_tf32_ts(td, tb + 8 * (i & 3), db, idesc, i > 0);
                else tc::mma_tf32_ss(td, da0 + (uint64_t)(i & 31), db, idesc, i > 0);
            }
            tc::commit(&bar);
        }
        tc::mbar_wait(&bar, parity); parity ^= 1;
        if (tid == 0) { long long dt = clock64() - t0; if (dt < best) best = dt; }
    }
    tc::fence_after();
    float v[16]; tc::tmem_ld16_nowait(td + lane_base, v); tc::tmem_wait_ld(v);
    sink[tid] = v[0];
    if (tid == 0) *cycles = best;
    tc::fence_before(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

// ---------------------------------------------------------------------------------------------------------------------------
// test 3: the production round trip, NWG independent warpgroups per CTA, grid = 148 x CTAs per SM
// ---------------------------------------------------------------------------------------------------------------------------
template <int KC, int NC, int NWG>
__global__ void __launch_bounds__(128 * NWG) k_rt(int iters, int do_mma, long long* cyc_out, float* sink) {
    extern __shared__ __align__(128) float dsm[];
    float* Bhi = dsm;                                 // [KC/4][NC][4]
    float* Blo = dsm + KC * NC;
    __shared__ __align__(8) uint64_t bars[NWG];
    __shared__ uint32_t tmem_base;
    constexpr uint32_t WG_COLS = 2 * KC + NC, COLS = pow2_cols(NWG * WG_COLS);
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, tig = tid & 127;
    for (int i = tid; i < KC * NC; i += 128 * NWG) { Bhi[i] = 0.01f * (float)((i * 7) % 13 - 6); Blo[i] = 1e-5f * (float)((i * 5) % 11 - 5); }
    if (warp == 0) tc::tmem_alloc(&tmem_base, COLS);
    if (tid < NWG) tc::mbar_init(&bars[tid], 1);
    tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t tb = tmem_base + wg * WG_COLS, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t A_HI = tb, A_LO = tb + KC, D0 = tb + 2 * KC;
    const uint32_t idesc = tc::make_idesc_tf32(128, NC);
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = 0.001f * (float)((tid * 16 + j) % 97);
    uint32_t parity = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < KC / 16; ++c) {
            float hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float v = x[j] + (float)c; hi[j] = tc::tf32_hi(v); lo[j] = v - hi[j]; }
            tc::tmem_st16(A_HI + lane_base + 16 * c, hi);
            tc::tmem_st16(A_LO + lane_base + 16 * c, lo);
        }
        tc::tmem_wait_st();
        tc::fence_before();
        tc::bar_sync(1 + wg, 128);
        if (tig == 0) {
            if (do_mma) {
                tc::fence_after();
                uint32_t acc = 0;
#pragma unroll
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t bh = tc::make_desc(tc::smem_u32(Bhi + 2 * ks * NC * 4), NC * 16, 128);
                    const uint64_t bl = tc::make_desc(tc::smem_u32(Blo + 2 * ks * NC * 4), NC * 16, 128);
                    tc::mma_tf32_ts(D0, A_LO + 8 * ks, bh, idesc, acc); acc = 1;
                    tc::mma_tf32_ts(D0, A_HI + 8 * ks, bl, idesc, 1);
                    tc::mma_tf32_ts(D0, A_HI + 8 * ks, bh, idesc, 1);
                }
                tc::commit(&bars[wg]);
            } else {
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(tc::smem_u32(&bars[wg])) : "memory");
            }
        }
        tc::mbar_wait(&bars[wg], parity); parity ^= 1;
        tc::fence_after();
#pragma unroll
        for (int c = 0; c < NC / 16; ++c) {
            float v[16];
            tc::tmem_ld16_nowait(D0 + lane_base + 16 * c, v);
            tc::tmem_wait_ld(v);
#pragma unroll
            for (int j = 0; j < 16; ++j) x[j] = fmaf(x[j], 0.25f, do_mma ? 1e-3f * v[j] : 0.f);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += x[j];
    sink[(size_t)blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0 && blockIdx.x == 0) *cyc_out = t1 - t0;
    tc::fence_before(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, COLS);
}


// ---------------------------------------------------------------------------------------------------------------------------
// test 4: kind::f16 with the A operand in TMEM as packed fp16 pairs: a 16-channel block is 8 columns of (hi[2c], hi[2c+1]) and
// 8 columns of (lo[2c], lo[2c+1]), hi = fp16(x), lo = fp16(x - hi) (11 + 11 mantissa bits, like the TF32 split, in half the TMEM
// columns and half the MMAs: K = 16 per instruction).  B = W_hi / W_lo fp16, K-major planes [k/8][n][8 halves].
// D[128][N] = sum over KB blocks of 16 channels.
// ---------------------------------------------------------------------------------------------------------------------------
template <int KB, int N>
__global__ void __launch_bounds__(128) k_f16_correct(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out, float wscale) {
    __shared__ __align__(128) __half Bhi[KB][2][N][8];
    __shared__ __align__(128) __half Blo[KB][2][N][8];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    constexpr uint32_t D0 = 16 * KB, COLS = pow2_cols(16 * KB + N);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < KB * 16 * N; i += 128) {
        const int k = i / N, n = i % N, kb = k >> 4, kk = k & 15;
        const float v = B[i] * wscale;
        const __half hi = __float2half_rn(v);
        Bhi[kb][kk >> 3][n][kk & 7] = hi;
        Blo[kb][kk >> 3][n][kk & 7] = __float2half_rn(v - __half2float(hi));
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base, COLS);
    if (tid == 0) tc::mbar_init(&bar, 1);
    tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t tb = tmem_base, lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float a = A[tid * 16 * KB + 16 * kb + 2 * c], b = A[tid * 16 * KB + 16 * kb + 2 * c + 1];
            hi[c] = tc::pack_f16x2(a, b);
            lo[c] = tc::pack_f16x2(a - tc::f16lo_to_f32(hi[c]), b - tc::f16hi_to_f32(hi[c]));
        }
        tc::tmem_st8u(tb + lane_base + 16 * kb, hi);
        tc::tmem_st8u(tb + lane_base + 16 * kb + 8, lo);
    }
    tc::tmem_wait_st(); tc::fence_before(); __syncthreads();
    if (tid == 0) {
        tc::fence_after();
        const uint32_t idesc = tc::make_idesc_f16(128, N);
        uint32_t acc = 0;
        for (int kb = 0; kb < KB; ++kb) {
            const uint64_t bh = tc::make_desc(tc::smem_u32(&Bhi[kb][0][0][0]), N * 16, 128);
            const uint64_t bl = tc::make_desc(tc::smem_u32(&Blo[kb][0][0][0]), N * 16, 128);
            tc::mma_f16_ts(tb + D0, tb + 16 * kb + 8, bh, idesc, acc); acc = 1;     // lo * W_hi
            tc::mma_f16_ts(tb + D0, tb + 16 * kb, bl, idesc, 1);                     // hi * W_lo
            tc::mma_f16_ts(tb + D0, tb + 16 * kb, bh, idesc, 1);                     // hi * W_hi
        }
        tc::commit(&bar);
    }
    tc::mbar_wait(&bar, 0); tc::fence_after();
#pragma unroll
    for (int c = 0; c < N / 16; ++c) {
        float v[16];
        tc::tmem_ld16_nowait(tb + lane_base + D0 + 16 * c, v);
        tc::tmem_wait_ld(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) out[tid * N + 16 * c + j] = v[j] / wscale;
    }
    tc::fence_before(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, COLS);
}

// test 5: MMA cost with NACC independent accumulators (round robin), TS form, N columns each
template <int N, int NACC>
__global__ void __launch_bounds__(128) k_time_acc(int nmma, int reps, long long* cycles, float* sink) {
    extern __shared__ __align__(128) float dsm[];
    float* B = dsm;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * N * 4; i += 128) B[i] = 0.5f;
    if (warp == 0) tc::tmem_alloc(&tmem_base, 512);
    if (tid == 0) tc::mbar_init(&bar, 1);
    tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t tb = tmem_base, lane_base = (uint32_t)(warp * 32) << 16;
    { float one[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) one[j] = 1.0f;
      tc::tmem_st16(tb + lane_base + 0, one); tc::tmem_st16(tb + lane_base + 16, one); tc::tmem_wait_st(); }
    tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t idesc = tc::make_idesc_tf32(128, N), td = tb + 64;
    const uint64_t db = tc::make_desc(tc::smem_u32(B), N * 16, 128);
    uint32_t parity = 0; long long best = 1ll << 60;
    for (int rep = 0; rep < reps; ++rep) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int i = 0; i < nmma; ++i) tc::mma_tf32_ts(td + N * (i % NACC), tb + 8 * (i & 3), db, idesc, i >= NACC);
            tc::commit(&bar);
        }
        tc::mbar_wait(&bar, parity); parity ^= 1;
        if (tid == 0) { long long dt = clock64() - t0; if (dt < best) best = dt; }
    }
    tc::fence_after();
    float v[16]; tc::tmem_ld16_nowait(td + lane_base, v); tc::tmem_wait_ld(v);
    sink[tid] = v[0];
    if (tid == 0) *cycles = best;
    tc::fence_before(); __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tb, 512);
}

template <int KB, int N>
static void run_f16(float amp) {
    std::vector<float> A(128 * 16 * KB), B(16 * KB * N), out(128 * N);
    float *dA, *dB, *dO;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
    for (auto& v : A) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * amp;
    float wmax = 0;
    for (auto& v : B) { v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f; wmax = fmaxf(wmax, fabsf(v)); }
    const float wscale = exp2f(floorf(log2f(16384.f / wmax)));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    k_f16_correct<KB, N><<<1, 128>>>(dA, dB, dO, wscale);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
    double es = 0, ef = 0, mx = 0;
    for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double rx = 0; float s = 0.f;
        for (int k = 0; k < 16 * KB; ++k) { rx += (double)A[r * 16 * KB + k] * B[k * N + n]; s = fmaf(A[r * 16 * KB + k], B[k * N + n], s); }
        es = fmax(es, fabs(rx - out[r * N + n])); ef = fmax(ef, fabs(rx - (double)s)); mx = fmax(mx, fabs(rx));
    }
    printf("test4 f16 hi/lo TS K=%d N=%d |a|<=%g wscale=%g: max err %.3e (rel %.2e, |ref|max %.2f); sequential fp32 FFMA: %.3e\n", 16 * KB, N, amp, wscale, es, es / mx, mx, ef);
    CK(cudaFree(dA)); CK(cudaFree(dB)); CK(cudaFree(dO));
}

template <int N, int NACC>
static void run_time_acc(long long* d_cyc, float* d_sink) {
    const size_t sm = (4 * N * 4) * 4;
    long long c[2]; int ns[2] = {24, 120};
    for (int i = 0; i < 2; ++i) { k_time_acc<N, NACC><<<1, 128, sm>>>(ns[i], 20, d_cyc, d_sink); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&c[i], d_cyc, 8, cudaMemcpyDeviceToHost)); }
    printf("test5 TS M128 N=%d K8, %d independent accumulators: 24/120 MMAs = %lld %lld cyc (%.1f per MMA)\n", N, NACC, c[0], c[1], (c[1] - c[0]) / 96.0);
}

template <int K, int N>
static void run_correct() {
    std::vector<float> A(128 * K), B(K * N), out(128 * N);
    float *dA, *dB, *dO;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dO, out.size() * 4));
    auto launch = [&](int mode) {
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset(dO, 0, out.size() * 4));
        k_ts_correct<K, N><<<1, 128>>>(dA, dB, dO, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost));
    };
    auto trunc = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    auto rnd = [](float v) { uint32_t u; memcpy(&u, &v, 4); u += 0x1000u; u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
    // (a) tf32-exact operands: must be exact
    for (auto& v : A) v = (float)((rand() % 33) - 16) / 8.0f;
    for (auto& v : B) v = (float)((rand() % 17) - 8) / 4.0f;
    launch(0);
    double e = 0;
    for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[r * K + k] * B[k * N + n];
        e = fmax(e, fabs(ref - out[r * N + n]));
    }
    printf("test1 TS K=%d N=%d tf32-exact operands : max err %.3e %s\n", K, N, e, e == 0 ? "EXACT" : "MISMATCH");
    // (b) arbitrary fp32 operands, one product: truncation or rounding of the low 13 bits?
    for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.3f;
    launch(0);
    double et = 0, er = 0, ex = 0, mx = 0;
    for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double rt = 0, rr = 0, rx = 0;
        for (int k = 0; k < K; ++k) {
            rt += (double)trunc(A[r * K + k]) * trunc(B[k * N + n]);
            rr += (double)rnd(A[r * K + k]) * rnd(B[k * N + n]);
            rx += (double)A[r * K + k] * B[k * N + n];
        }
        et = fmax(et, fabs(rt - out[r * N + n])); er = fmax(er, fabs(rr - out[r * N + n])); ex = fmax(ex, fabs(rx - out[r * N + n]));
        mx = fmax(mx, fabs(rx));
    }
    printf("test1 TS K=%d N=%d fp32 operands, 1 product: err vs truncated-operand model %.3e, vs rounded-operand model %.3e, vs exact %.3e (|ref|max %.2f)\n",
           K, N, et, er, ex, mx);
    // (c) 3xTF32 split
    launch(1);
    double es = 0, ef = 0;
    for (int r = 0; r < 128; ++r) for (int n = 0; n < N; ++n) {
        double rx = 0; float s = 0.f;
        for (int k = 0; k < K; ++k) { rx += (double)A[r * K + k] * B[k * N + n]; s = fmaf(A[r * K + k], B[k * N + n], s); }
        es = fmax(es, fabs(rx - out[r * N + n])); ef = fmax(ef, fabs(rx - (double)s));
    }
    printf("test1 TS K=%d N=%d 3xTF32 split            : max err %.3e (rel %.2e); sequential fp32 FFMA of the same sums: %.3e\n", K, N, es, es / mx, ef);
    CK(cudaFree(dA)); CK(cudaFree(dB)); CK(cudaFree(dO));
}

template <int N>
static void run_time(long long* d_cyc, float* d_sink) {
    const size_t sm = (4 * 160 * 4 + 4 * N * 4) * 4;
    CK(cudaFuncSetAttribute(k_time<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(k_time<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    long long c[2][3]; int ns[3] = {1, 16, 128};
    for (int ts = 0; ts < 2; ++ts)
        for (int i = 0; i < 3; ++i) {
            if (ts) k_time<N, true><<<1, 128, sm>>>(ns[i], 20, d_cyc, d_sink); else k_time<N, false><<<1, 128, sm>>>(ns[i], 20, d_cyc, d_sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&c[ts][i], d_cyc, 8, cudaMemcpyDeviceToHost));
        }
    printf("test2 M128 N=%3d K8: SS 1/16/128 MMAs = %5lld %6lld %7lld cyc (%.1f per MMA)   TS = %5lld %6lld %7lld cyc (%.1f per MMA)\n", N,
           c[0][0], c[0][1], c[0][2], (c[0][2] - c[0][1]) / 112.0, c[1][0], c[1][1], c[1][2], (c[1][2] - c[1][1]) / 112.0);
}

template <int KC, int NC, int NWG>
static void run_rt(int ctas_per_sm, int iters, long long* d_cyc, float* d_sink) {
    // dynamic shared memory sized so that exactly ctas_per_sm CTAs fit on an SM
    size_t sm = (size_t)(220 * 1024) / ctas_per_sm - 2048;
    const size_t need = (size_t)2 * KC * NC * 4;
    if (sm < need) { printf("test3 K=%d N=%d NWG=%d x %d CTAs: not enough shared memory\n", KC, NC, NWG, ctas_per_sm); return; }
    CK(cudaFuncSetAttribute(k_rt<KC, NC, NWG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int mma = 1; mma >= 0; --mma) {
        k_rt<KC, NC, NWG><<<148 * ctas_per_sm, 128 * NWG, sm>>>(50, mma, d_cyc, d_sink);     // warm-up
        CK(cudaEventRecord(e0));
        k_rt<KC, NC, NWG><<<148 * ctas_per_sm, 128 * NWG, sm>>>(iters, mma, d_cyc, d_sink);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        const double tiles_per_sm = (double)iters * NWG * ctas_per_sm;
        printf("test3 K=%2d N=%2d  %d WG/CTA x %d CTA/SM  %s: %7.1f cyc per round trip (one WG), %6.1f cyc per 128-row tile per SM, kernel %.3f ms\n",
               KC, NC, NWG, ctas_per_sm, mma ? "st+MMA+ld" : "st+ld only", (double)c / iters, (double)c / tiles_per_sm, ms);
    }
}

int main() {
    srand(3);
    run_correct<16, 16>();
    run_correct<48, 48>();
    long long* d_cyc; float* d_sink;
    CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_sink, 148 * 8 * 1024 * 4));
    run_f16<1, 16>(1.f); run_f16<3, 48>(1.f); run_f16<3, 48>(100.f); run_f16<3, 48>(0.001f);
    run_time_acc<16, 1>(d_cyc, d_sink); run_time_acc<16, 2>(d_cyc, d_sink); run_time_acc<16, 3>(d_cyc, d_sink); run_time_acc<16, 6>(d_cyc, d_sink);
    run_time_acc<48, 1>(d_cyc, d_sink); run_time_acc<48, 2>(d_cyc, d_sink); run_time_acc<48, 3>(d_cyc, d_sink);
    if (getenv("UMMA_SHORT")) { printf("done\n"); return 0; }
    run_time<16>(d_cyc, d_sink); run_time<48>(d_cyc, d_sink); run_time<64>(d_cyc, d_sink);
    run_time<128>(d_cyc, d_sink); run_time<144>(d_cyc, d_sink); run_time<256>(d_cyc, d_sink);
    const int it = 2000;
    // GTCN pattern: K = 16 -> N = 16 pointwise conv, 48 TMEM columns per warpgroup
    run_rt<16, 16, 1>(1, it, d_cyc, d_sink); run_rt<16, 16, 1>(2, it, d_cyc, d_sink); run_rt<16, 16, 1>(4, it, d_cyc, d_sink); run_rt<16, 16, 1>(8, it, d_cyc, d_sink);
    run_rt<16, 16, 2>(4, it, d_cyc, d_sink); run_rt<16, 16, 4>(2, it, d_cyc, d_sink); run_rt<16, 16, 8>(1, it, d_cyc, d_sink);
    // decoder 3x3 pattern: K = 48 (3 frames x 16 ci) -> N = 48 (3 kf x 16 co), 144 TMEM columns per warpgroup
    run_rt<48, 48, 1>(1, it, d_cyc, d_sink); run_rt<48, 48, 1>(2, it, d_cyc, d_sink); run_rt<48, 48, 3>(1, it, d_cyc, d_sink);
    // scatter form: K = 16 -> N = 144
    run_rt<16, 144, 1>(1, it, d_cyc, d_sink); run_rt<16, 144, 1>(2, it, d_cyc, d_sink);
    printf("done\n");
    return 0;
}
