// Prototype / unit test for the tcgen05 (UMMA) TF32 path planned for the GTConv 3x3 and 1x1 contractions.
//
// Validates on real hardware, before the production kernels depend on it:
//   1. the no-swizzle K-major shared-memory descriptors for activation "planes" [ci/4][row][4 floats]
//      (rows 16 B apart, SBO = 128 B, LBO = plane stride) including ROW-SHIFTED start addresses -- a conv tap
//      (kt,kf) is just a different start row of the same plane, so a 3x3 conv is 9 taps x (K/8) MMAs;
//   2. weights as planes [ci/4][co 0..15][4 floats] (N = 16, K-major);
//   3. the 3xTF32 split (hi*hi + lo*hi + hi*lo) reaching fp32-class accuracy;
//   4. M=128, N=16, K=8 instruction throughput.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_tf32_test umma_tf32_test.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

namespace um {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE: element (row r, 16-byte k-chunk j) at start + (r%8)*16 + (r/8)*SBO + j*LBO
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // descriptor version 1 (sm_100)
    return d;                                     // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// kind::tf32, D = f32, A/B K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra WAIT_LOOP;\n\t}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
}  // namespace um

constexpr int ROWS = 128 + 80;       // activation rows available (row-shifted taps read up to 128 + shift)
constexpr int NTAP = 9;

// mode 0: single tap, K=8, no split (inputs must be tf32-exact)   mode 1: 9 row-shifted taps, K=16, 3xTF32 split
__global__ void __launch_bounds__(128) k_test(const float* __restrict__ act /*[ROWS][16]*/, const float* __restrict__ w /*[NTAP][16 ci][16 co]*/,
                                              const int* __restrict__ shifts /*[NTAP] row shift per tap*/, float* __restrict__ out /*[128][16]*/,
                                              int mode, int reps, long long* cycles) {
    __shared__ __align__(128) float Ahi[4][ROWS][4];
    __shared__ __align__(128) float Alo[4][ROWS][4];
    __shared__ __align__(128) float Bhi[NTAP][4][16][4];
    __shared__ __align__(128) float Blo[NTAP][4][16][4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < ROWS * 16; i += 128) {
        const int r = i / 16, c = i % 16;
        const float v = act[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        Ahi[c / 4][r][c % 4] = hi;
        Alo[c / 4][r][c % 4] = v - hi;
    }
    for (int i = tid; i < NTAP * 256; i += 128) {
        const int t = i / 256, ci = (i / 16) % 16, co = i % 16;
        const float v = w[i];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        Bhi[t][ci / 4][co][ci % 4] = hi;
        Blo[t][ci / 4][co][ci % 4] = v - hi;
    }
    if (warp == 0) um::tmem_alloc(&tmem_base, 32);
    if (tid == 0) um::mbar_init(&bar, 1);
    um::fence_async_smem();                       // generic-proxy smem writes -> visible to the async (tensor core) proxy
    um::tc_fence_before();
    __syncthreads();
    um::tc_fence_after();
    const uint32_t td = tmem_base;
    const uint32_t idesc = um::make_idesc_tf32(128, 16);
    const uint32_t plane_bytes = ROWS * 16;       // LBO: next 16-byte k-chunk of the same row
    long long t0 = 0, t1 = 0;
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            t0 = clock64();
            if (mode == 0) {
                const uint64_t da = um::make_desc(um::smem_u32(&Ahi[0][shifts[0]][0]), plane_bytes, 128);
                const uint64_t db = um::make_desc(um::smem_u32(&Bhi[0][0][0][0]), 256, 128);
                um::mma_tf32(td, da, db, idesc, 0);
            } else {
                uint32_t acc = 0;
                for (int t = 0; t < NTAP; ++t) {
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t ah = um::make_desc(um::smem_u32(&Ahi[2 * ks][shifts[t]][0]), plane_bytes, 128);
                        const uint64_t al = um::make_desc(um::smem_u32(&Alo[2 * ks][shifts[t]][0]), plane_bytes, 128);
                        const uint64_t bh = um::make_desc(um::smem_u32(&Bhi[t][2 * ks][0][0]), 256, 128);
                        const uint64_t bl = um::make_desc(um::smem_u32(&Blo[t][2 * ks][0][0]), 256, 128);
                        if (mode == 1) {
                            um::mma_tf32(td, al, bh, idesc, acc); acc = 1;      // small terms first
                            um::mma_tf32(td, ah, bl, idesc, 1);
                        }
                        um::mma_tf32(td, ah, bh, idesc, acc); acc = 1;
                    }
                }
            }
            um::commit(&bar);
        }
        um::mbar_wait(&bar, parity);
        parity ^= 1;
        if (tid == 0) t1 = clock64();
    }
    um::tc_fence_after();
    float v[16];
    um::tmem_ld16(td + ((uint32_t)(warp * 32) << 16), v);      // lane field = bits 31:16
#pragma unroll
    for (int c = 0; c < 16; ++c) out[tid * 16 + c] = v[c];
    if (tid == 0 && cycles) *cycles = t1 - t0;
    um::tc_fence_before();
    __syncthreads();
    if (warp == 0) um::tmem_dealloc(td, 32);
}

// timing: NMMA back-to-back MMAs with precomputed descriptors, N columns, then commit + wait
template <int N>
__global__ void __launch_bounds__(128) k_time(int nmma, int reps, long long* cycles, float* sink) {
    __shared__ __align__(128) float A[4][ROWS][4];
    __shared__ __align__(128) float B[4][64][4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 4 * ROWS * 4; i += 128) (&A[0][0][0])[i] = 1.0f;
    for (int i = tid; i < 4 * 64 * 4; i += 128) (&B[0][0][0])[i] = 0.5f;
    if (warp == 0) um::tmem_alloc(&tmem_base, 64);
    if (tid == 0) um::mbar_init(&bar, 1);
    um::fence_async_smem(); um::tc_fence_before(); __syncthreads(); um::tc_fence_after();
    const uint32_t td = tmem_base, idesc = um::make_idesc_tf32(128, N);
    const uint64_t da0 = um::make_desc(um::smem_u32(&A[0][0][0]), ROWS * 16, 128);
    const uint64_t db = um::make_desc(um::smem_u32(&B[0][0][0]), 64 * 16, 128);
    uint32_t parity = 0; long long best = 1ll << 60;
    for (int rep = 0; rep < reps; ++rep) {
        long long t0 = 0;
        if (tid == 0) {
            t0 = clock64();
            for (int i = 0; i < nmma; ++i) um::mma_tf32(td, da0 + (uint64_t)(i & 31), db, idesc, i > 0);   // start row shifts 0..31
            um::commit(&bar);
        }
        um::mbar_wait(&bar, parity); parity ^= 1;
        if (tid == 0) { long long dt = clock64() - t0; if (dt < best) best = dt; }
    }
    um::tc_fence_after();
    float v[16]; um::tmem_ld16(td + ((uint32_t)(warp * 32) << 16), v);
    sink[tid] = v[0];
    if (tid == 0) *cycles = best;
    um::tc_fence_before(); __syncthreads();
    if (warp == 0) um::tmem_dealloc(td, 64);
}

int main() {
    std::vector<float> act(ROWS * 16), w(NTAP * 256), out(128 * 16);
    std::vector<int> shifts(NTAP);
    float *d_act, *d_w, *d_out; int* d_sh; long long* d_cyc;
    CK(cudaMalloc(&d_act, act.size() * 4)); CK(cudaMalloc(&d_w, w.size() * 4)); CK(cudaMalloc(&d_out, out.size() * 4));
    CK(cudaMalloc(&d_sh, NTAP * 4)); CK(cudaMalloc(&d_cyc, 8));
    srand(3);
    auto run = [&](int mode, int reps) {
        CK(cudaMemcpy(d_act, act.data(), act.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_w, w.data(), w.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_sh, shifts.data(), NTAP * 4, cudaMemcpyHostToDevice));
        k_test<<<1, 128>>>(d_act, d_w, d_sh, d_out, mode, reps, d_cyc);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        return c;
    };
    // ---- test 1: one MMA, K = 8, tf32-exact inputs, row shifts 0 / 1 / 37
    for (int sh : {0, 1, 37}) {
        for (auto& v : act) v = (float)((rand() % 33) - 16) / 8.0f;
        for (auto& v : w) v = (float)((rand() % 17) - 8) / 4.0f;
        shifts.assign(NTAP, 0); shifts[0] = sh;
        run(0, 1);
        double maxerr = 0;
        for (int r = 0; r < 128; ++r)
            for (int co = 0; co < 16; ++co) {
                double ref = 0;
                for (int ci = 0; ci < 8; ++ci) ref += (double)act[(r + sh) * 16 + ci] * w[ci * 16 + co];
                maxerr = fmax(maxerr, fabs(ref - out[r * 16 + co]));
            }
        printf("test1 single MMA K=8 shift=%2d : max err %.3e %s\n", sh, maxerr, maxerr == 0 ? "EXACT" : "MISMATCH");
    }
    // ---- test 2: 3x3-conv-like: 9 row-shifted taps (kt*35 + kf), K = 16, random fp32 inputs
    for (auto& v : act) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : w) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.2f;
    for (int kt = 0; kt < 3; ++kt) for (int kf = 0; kf < 3; ++kf) shifts[kt * 3 + kf] = kt * 35 + kf;
    std::vector<double> ref(128 * 16, 0.0);
    for (int r = 0; r < 128; ++r)
        for (int co = 0; co < 16; ++co) {
            double s = 0;
            for (int t = 0; t < NTAP; ++t)
                for (int ci = 0; ci < 16; ++ci) s += (double)act[(r + shifts[t]) * 16 + ci] * w[t * 256 + ci * 16 + co];
            ref[r * 16 + co] = s;
        }
    for (int mode : {2, 1}) {
        run(mode, 1);
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < 128 * 16; ++i) { maxerr = fmax(maxerr, fabs(ref[i] - out[i])); maxref = fmax(maxref, fabs(ref[i])); }
        printf("test2 9 taps K=16 %-14s : max err %.3e (|ref|max %.2f, rel %.2e)\n", mode == 1 ? "3xTF32 split" : "plain TF32(hi)", maxerr, maxref, maxerr / maxref);
    }
    // fp32 FFMA-order reference error for comparison
    {
        double maxerr = 0;
        for (int r = 0; r < 128; ++r)
            for (int co = 0; co < 16; ++co) {
                float s = 0;
                for (int t = 0; t < NTAP; ++t)
                    for (int ci = 0; ci < 16; ++ci) s = fmaf(act[(r + shifts[t]) * 16 + ci], w[t * 256 + ci * 16 + co], s);
                maxerr = fmax(maxerr, fabs(ref[r * 16 + co] - s));
            }
        printf("      (sequential fp32 FFMA accumulation of the same sum: max err %.3e)\n", maxerr);
    }
    // ---- timing: 54 MMAs (9 taps x 2 k-steps x 3 split terms) + commit, one CTA
    long long c = run(1, 200);
    printf("timing: 54 x (M128 N16 K8) tf32 MMAs + commit + wait = %lld cycles  (%.1f cycles per MMA)\n", c, c / 54.0);
    c = run(2, 200);
    printf("timing: 18 x (M128 N16 K8) tf32 MMAs + commit + wait = %lld cycles  (%.1f cycles per MMA)\n", c, c / 18.0);
    {
        float* d_sink; CK(cudaMalloc(&d_sink, 512));
        auto tm = [&](auto kern, int n) { kern<<<1, 128>>>(n, 20, d_cyc, d_sink); CK(cudaDeviceSynchronize()); long long c2; CK(cudaMemcpy(&c2, d_cyc, 8, cudaMemcpyDeviceToHost)); return c2; };
        for (int n : {1, 2, 8, 32, 128}) {
            long long c16 = tm(k_time<16>, n), c32 = tm(k_time<32>, n), c64 = tm(k_time<64>, n);
            printf("timing precomputed desc: %3d MMAs (M128 K8): N=16 %6lld cyc  N=32 %6lld cyc  N=64 %6lld cyc\n", n, c16, c32, c64);
        }
    }
    printf("done\n");
    return 0;
}
