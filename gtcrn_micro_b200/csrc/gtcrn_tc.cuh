// Tensor-core (tcgen05) versions of the two FMA-bound layer types of GTCRN-Micro: the GTCN pointwise convolutions and the
// decoder GTConv 3x3 convolution (63 % of the model's MACs).  fp32-class accuracy through the 3xTF32 split.
//
// Why this works where round 1's attempt did not: that prototype fed A from shared memory, and an
// M128 x N<=64 x K8 MMA then costs ~55 cycles whatever N is -- no better than FFMA for N = 16.  Here a thread owns one position
// (one row of the M = 128 tile) and already holds that row's activations in registers, which is exactly the tensor-memory layout
// of an A operand: row m in TMEM lane m, element k in column k.  tcgen05.st writes A straight from registers, tcgen05.mma reads A
// from TMEM and the (tiny, constant) weight matrix B from shared memory, tcgen05.ld brings the fp32 accumulator row back to the
// same thread for bias / PReLU / the depthwise taps.  Shared memory carries no activations for the contractions at all.
// Measured on B200 (tools/microbench/umma_ts_test.cu, profiles/umma_ts_test_r2.txt): a dependent MMA costs ~46 cycles for
// N <= 96 and N/2 cycles above, MMAs of different CTAs overlap, and with 7-8 warpgroups per SM a whole
// split -> st -> MMA -> ld round trip sustains one 128-row tile per ~150 cycles per SM.
//
// 3xTF32: x = hi + lo, hi = x truncated to TF32 (the tensor core ignores the low 13 mantissa bits: measured), lo = x - hi exact;
// x*w ~ hi*w_hi + lo*w_hi + hi*w_lo.  The weights are split once per CTA, rounded to nearest, and stored as [W_hi | W_lo] side by
// side along N, so one MMA with A_hi yields hi*w_hi and hi*w_lo in two column blocks of the accumulator and a second, narrower
// MMA adds lo*w_hi: 2 MMAs per K = 8 step instead of 3; the two column blocks are added in registers after tcgen05.ld.
#pragma once
#include <algorithm>
#include <type_traits>
#include "gtcrn_kernels.cuh"
#include "gtcrn_stream.cuh"
#include <cuda_fp16.h>
#include "tc_util.cuh"

namespace gk {

// round-to-nearest TF32 split of a weight (done once per CTA): hi has 10 mantissa bits, lo = w - hi is rounded to TF32 as well
__device__ __forceinline__ void tf32_split_rn(float w, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u);
    lo = __uint_as_float((__float_as_uint(w - hi) + 0x1000u) & 0xFFFFE000u);
}

// ================================================================================================
// GTCN on tensor cores.  Same walking scheme as k_gtcn (one warp per (utterance, bin) column, lanes = 32 consecutive frames, x in
// registers through the four TCNs, conv1 outputs exchanged through the warp's private shared memory for the dilated depthwise
// taps), but the two 16 -> 16 pointwise convolutions of every TCN run as M128 x N32/16 x K8 tcgen05 MMAs over the 4 warps
// (= 4 columns = 128 rows) of a warpgroup.  reference: gtcrn_micro.py:256-336
//   TMEM per warpgroup (64 columns): D[0:32] accumulator (hi*w_hi | hi*w_lo), A_hi[32:48], A_lo[48:64]
//   shared memory: B = 8 matrices (4 TCNs x conv1 / conv3) x [4 k-planes][32 rows = W_hi co 0..15 | W_lo co 0..15][4] (16 KB),
//                  per warp Y[4 quads][32 frames] + H[4 quads][30 history rows] float4 (3.9 KB)
// ================================================================================================
constexpr int TCT_B_FLOATS = 8 * 512;
constexpr int TCT_WARP4 = 4 * 32 + 4 * 30;                 // float4 per warp
__host__ __device__ constexpr int tct_smem_floats(int nwg) { return TCT_B_FLOATS + nwg * 4 * TCT_WARP4 * 4; }
__host__ __device__ constexpr int tct_min_blocks(int nwg) { return nwg == 1 ? 6 : nwg == 2 ? 3 : 1; }   // shared memory allows 6 warpgroups per SM

// one pointwise 16 -> 16 convolution of the warpgroup's 128 rows, first half: v (registers) -> A in TMEM, MMAs, completion wait.
// STACK: 4 MMAs, the accumulator holds hi*w_hi + lo*w_hi in columns 0..15 and hi*w_lo in 16..31 (dependent chain of 4);
// otherwise 6 MMAs accumulate all three terms into columns 0..15 (chain of 6, half the tcgen05.ld traffic).
template <bool STACK, bool WAIT = true>
__device__ __forceinline__ void tct_conv16_issue(const float (&v)[16], uint32_t tm, uint32_t lane_base, uint64_t bdesc, uint64_t* bar,
                                                  uint32_t& parity, int bar_id, bool issuer) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                  // in halves: 16 instead of 32 staging registers
        float hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; c += 2) {
            hi[c] = tc::tf32_hi(v[8 * h + c]); hi[c + 1] = tc::tf32_hi(v[8 * h + c + 1]);
            lo[c] = v[8 * h + c]; lo[c + 1] = v[8 * h + c + 1];
            add_pair(lo[c], lo[c + 1], -hi[c], -hi[c + 1]);
        }
        tc::tmem_st8(tm + lane_base + 32 + 8 * h, hi);
        tc::tmem_st8(tm + lane_base + 48 + 8 * h, lo);
    }
    tc::tmem_wait_st();
    tc::fence_before();
    tc::bar_sync(bar_id, 128);
    if (issuer) {
        tc::fence_after();
        constexpr uint32_t id32 = tc::make_idesc_tf32(128, 32), id16 = tc::make_idesc_tf32(128, 16);
        if (STACK) {
            tc::mma_tf32_ts(tm, tm + 32, bdesc, id32, 0);              // k 0..7  : hi * [W_hi | W_lo]
            tc::mma_tf32_ts(tm, tm + 48, bdesc, id16, 1);              //           lo * W_hi
            tc::mma_tf32_ts(tm, tm + 40, bdesc + 64, id32, 1);         // k 8..15 (two k-planes = 1024 B further)
            tc::mma_tf32_ts(tm, tm + 56, bdesc + 64, id16, 1);
        } else {
            tc::mma_tf32_ts(tm, tm + 48, bdesc, id16, 0);              // small terms first: lo * W_hi
            tc::mma_tf32_ts(tm, tm + 32, bdesc + 16, id16, 1);         // hi * W_lo (rows 16..31 of the planes = 256 B further)
            tc::mma_tf32_ts(tm, tm + 56, bdesc + 64, id16, 1);
            tc::mma_tf32_ts(tm, tm + 40, bdesc + 64 + 16, id16, 1);
            tc::mma_tf32_ts(tm, tm + 32, bdesc, id16, 1);              // hi * W_hi
            tc::mma_tf32_ts(tm, tm + 40, bdesc + 64, id16, 1);
        }
        tc::commit(bar);
    }
    if (WAIT) {
        tc::mbar_wait(bar, parity);
        parity ^= 1;
        tc::fence_after();
    }
}
// completion wait of a convolution started with WAIT = false (whatever does not need its result goes in between)
__device__ __forceinline__ void tct_conv16_wait(uint64_t* bar, uint32_t& parity) {
    tc::mbar_wait(bar, parity);
    parity ^= 1;
    tc::fence_after();
}
// second half: accumulator columns 8h .. 8h+7 of this thread's row
template <bool STACK>
__device__ __forceinline__ void tct_fetch8(uint32_t tm, uint32_t lane_base, int h, float (&r)[8]) {
    tc::tmem_ld8_nowait(tm + lane_base + 8 * h, r);
    if (STACK) {
        float r1[8];
        tc::tmem_ld8_nowait(tm + lane_base + 16 + 8 * h, r1);
        tc::tmem_wait_ld(r);
        tc::tmem_wait_ld(r1);
#pragma unroll
        for (int c = 0; c < 8; c += 2) add_pair(r[c], r[c + 1], r1[c], r1[c + 1]);
    } else {
        tc::tmem_wait_ld(r);
    }
}

#ifdef GTCRNB_FB_TRACE
__device__ long long g_tcn_trace[4][16];
#define TCT(i) do { if (blockIdx.x == 3000 && blockIdx.y == 0 && lane == 0 && k == 2 && t0 == tw0 + 96) { long long c_; \
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)); g_tcn_trace[warp & 3][i] = c_; } } while (0)
#else
#define TCT(i) do { } while (0)
#endif
template <int NWG, bool STACK>
__global__ void __launch_bounds__(128 * NWG, tct_min_blocks(NWG))
k_gtcn_tc(const float* __restrict__ xin, float* __restrict__ out, const __grid_constant__ ParamBlock<4 * blob::TCN_SIZE> prm,
          int T, int seg_len, int64_t ncols) {
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) uint64_t bars[NWG];
    __shared__ uint32_t tmem_base_s;
    float* Bw = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wg = tid >> 7;
    float4* Y = smem4 + TCT_B_FLOATS / 4 + warp * TCT_WARP4;
    float4* H = Y + 4 * 32;

    // ---- prologue: split weights into [W_hi | W_lo] k-planes, allocate tensor memory
    for (int i = tid; i < 8 * 256; i += 128 * NWG) {
        const int m = i >> 8, ci = (i >> 4) & 15, co = i & 15;
        const float w = prm.w[(m >> 1) * blob::TCN_SIZE + ((m & 1) ? blob::TCN_W3 : blob::TCN_W1) + ci * 16 + co];
        float hi, lo;
        tf32_split_rn(w, hi, lo);
        float* pl = Bw + m * 512 + (ci >> 2) * 128 + (ci & 3);
        pl[co * 4] = hi;
        pl[(16 + co) * 4] = lo;
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 64 * NWG);
    if (tid < NWG) tc::mbar_init(&bars[tid], 1);
    for (int i = lane; i < 120; i += 32) H[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tm = tmem_base_s + 64 * wg, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint64_t bdesc0 = tc::make_desc(tc::smem_u32(Bw), 512, 128);
    const bool issuer = (tid & 127) == 0;
    uint64_t* bar = &bars[wg];
    uint32_t parity = 0;

    // every warp of a warpgroup takes part in every MMA round: out-of-range columns recompute the last valid one and do not store
    const int64_t col_raw = (int64_t)blockIdx.x * (4 * NWG) + warp;
    const bool col_ok = col_raw < ncols;
    const int64_t colid = col_ok ? col_raw : ncols - 1;
    const int64_t b = colid / kF2;
    const int f = (int)(colid - b * kF2);
    const int s0 = blockIdx.y * seg_len, s1 = min(T, s0 + seg_len);
    if (s0 < s1) {
        const int tw0 = max(0, s0 - TC_WARMUP);
        const float* xcol = xin + (b * T * kF2 + f) * 16;
        float* ocol = out + (b * T * kF2 + f) * 16;

        float4 xn[4];                                                 // rows of the next tile (loaded one tile ahead)
        {
            const int t = tw0 + lane;
            if (t < s1) {
                const float4* px = reinterpret_cast<const float4*>(xcol + (int64_t)t * (kF2 * 16));
#pragma unroll
                for (int v = 0; v < 4; ++v) xn[v] = px[v];
                if (t + 32 < s1) asm volatile("prefetch.global.L2 [%0];" ::"l"(xcol + (int64_t)(t + 32) * (kF2 * 16)));
            } else {
#pragma unroll
                for (int v = 0; v < 4; ++v) xn[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        for (int t0 = tw0; t0 < s1; t0 += 32) {
            const int t = t0 + lane;
            float x[16];
#pragma unroll
            for (int v = 0; v < 4; ++v) { x[4 * v] = xn[v].x; x[4 * v + 1] = xn[v].y; x[4 * v + 2] = xn[v].z; x[4 * v + 3] = xn[v].w; }
            auto tcn = [&](const int k, auto last_c) {

                const int d = 1 << k, coff = (2 << k) - 2, wo = k * blob::TCN_SIZE;
                float y[16];
                TCT(0);
                tct_conv16_issue<STACK>(x, tm, lane_base, bdesc0 + (uint64_t)(2 * k) * 128, bar, parity, 1 + wg, issuer);
                TCT(1);
                {
                    const float a1 = prm.w[wo + blob::TCN_SLOPE + 0];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float r[8];
                        tct_fetch8<STACK>(tm, lane_base, h, r);
#pragma unroll
                        for (int c = 0; c < 8; c += 2) {
                            add_pair(r[c], r[c + 1], prm.w[wo + blob::TCN_B1 + 8 * h + c], prm.w[wo + blob::TCN_B1 + 8 * h + c + 1]);
                            prelu_pair(r[c], r[c + 1], a1);
                            y[8 * h + c] = r[c]; y[8 * h + c + 1] = r[c + 1];
                        }
                    }
                }
                TCT(2);
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) Y[qd * 32 + lane] = make_float4(y[4 * qd], y[4 * qd + 1], y[4 * qd + 2], y[4 * qd + 3]);
                __syncwarp();
                TCT(3);
                {
                    // depthwise dilated taps, one channel quad at a time (8 instead of 32 tap registers live)
                    const float a2 = prm.w[wo + blob::TCN_SLOPE + 1];
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float4 ta = lane >= 2 * d ? Y[qd * 32 + lane - 2 * d] : H[qd * 30 + coff + lane];
                        const float4 tb = lane >= d ? Y[qd * 32 + lane - d] : H[qd * 30 + coff + lane + d];
#pragma unroll
                        for (int i = 0; i < 4; i += 2) {
                            const int c = 4 * qd + i;
                            float v0 = prm.w[wo + blob::TCN_B2 + c], v1 = prm.w[wo + blob::TCN_B2 + c + 1];
                            fma_pair2(v0, v1, f4get(ta, i), f4get(ta, i + 1), prm.w[wo + blob::TCN_W2 + c], prm.w[wo + blob::TCN_W2 + c + 1]);
                            fma_pair2(v0, v1, f4get(tb, i), f4get(tb, i + 1), prm.w[wo + blob::TCN_W2 + 16 + c], prm.w[wo + blob::TCN_W2 + 16 + c + 1]);
                            fma_pair2(v0, v1, y[c], y[c + 1], prm.w[wo + blob::TCN_W2 + 32 + c], prm.w[wo + blob::TCN_W2 + 32 + c + 1]);
                            prelu_pair(v0, v1, a2);
                            y[c] = v0; y[c + 1] = v1;
                        }
                    }
                }
                if constexpr (decltype(last_c)::value) {
                    // next tile's rows, requested before the LAST convolution of this tile (an L2 hit costs > 2,000 cycles under
                    // this load; at the top of the next tile nothing covered it).  Only this peeled copy of the TCN body carries them.
                    const int tn = t + 32;
                    if (tn < s1) {
                        const float4* pn = reinterpret_cast<const float4*>(xcol + (int64_t)tn * (kF2 * 16));
#pragma unroll
                        for (int v = 0; v < 4; ++v) xn[v] = pn[v];
                        if (tn + 32 < s1) asm volatile("prefetch.global.L2 [%0];" ::"l"(xcol + (int64_t)(tn + 32) * (kF2 * 16)));
                    } else {
#pragma unroll
                        for (int v = 0; v < 4; ++v) xn[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                TCT(4);
                tct_conv16_issue<STACK, false>(y, tm, lane_base, bdesc0 + (uint64_t)(2 * k + 1) * 128, bar, parity, 1 + wg, issuer);
                TCT(5);
                // history <- the last 2d conv1 rows of this tile, WHILE the MMAs of the second convolution run (the clock64 trace showed
                // this copy at ~550 of a TCN's 4,450 cycles after the completion wait).  The warpgroup barrier inside the call above
                // separates these writes from every lane's tap reads, the one of the next conv1 separates the Y reads from the next Y writes.
                if (lane < 8 * d) {                                         // (shifts and masks: 2d = 2 << k)
                    const int qd = lane >> (k + 1), j = lane & (2 * d - 1);
                    H[qd * 30 + coff + j] = Y[qd * 32 + 32 - 2 * d + j];
                }
                if (k == 3) {                                               // d = 8: 64 float4 to move
                    const int i = lane + 32, qd = i >> 4, j = i & 15;
                    H[qd * 30 + coff + j] = Y[qd * 32 + 16 + j];
                }
                tct_conv16_wait(bar, parity);
                TCT(6);
                {
                    const float a3 = prm.w[wo + blob::TCN_SLOPE + 2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float r[8];
                        tct_fetch8<STACK>(tm, lane_base, h, r);
#pragma unroll
                        for (int c = 0; c < 8; c += 2) {
                            add_pair(r[c], r[c + 1], prm.w[wo + blob::TCN_B3 + 8 * h + c], prm.w[wo + blob::TCN_B3 + 8 * h + c + 1]);
                            add_pair(x[8 * h + c], x[8 * h + c + 1], r[c], r[c + 1]);
                            prelu_pair(x[8 * h + c], x[8 * h + c + 1], a3);
                        }
                    }
                }
                TCT(7);
                        };
#pragma unroll 1
            for (int k = 0; k < 3; ++k) tcn(k, std::false_type{});
            tcn(3, std::true_type{});                                   // peeled: carries the next tile's loads
            if (col_ok && t >= s0 && t < s1) {
                float4* po = reinterpret_cast<float4*>(ocol + (int64_t)t * (kF2 * 16));
#pragma unroll
                for (int v = 0; v < 4; ++v) po[v] = make_float4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 64 * NWG);
}

// ================================================================================================
// Decoder GTConvBlock on tensor cores.  reference: gtcrn_micro.py:176-254 (GTConvBlock), :441-460 (decoder use), TRALite :112-174
//
// A warpgroup owns GTT_G = 3 (utterance, segment) groups x 33 bins = 99 of the 128 rows of an M = 128 tile and walks along
// time ONE FRAME PER STEP; thread (g, f) is row g*33 + f.  The full 16 -> 16 3x3 convolution (2,304 of the block's 2,560 MACs per
// position) is one accumulator D[128 x 48] per step:
//      D_f[kf*16 + co] = sum_{kt, ci} p[t-2+kt][f][ci] * W[kt][kf][ci][co]                     (K = 3 frames x 16 ci = 48)
//      m[t][f]         = D_{f-1}[kf=0] + D_f[kf=1] + D_{f+1}[kf=2] + b                          (zero padded in f)
// The three K blocks are the point-conv outputs p of frames t-2, t-1, t AT THE SAME BIN, i.e. values the same thread produced in
// its last three steps: they sit in a 3-slot ring of TMEM columns, every step overwrites the oldest slot with the new frame and
// pairs each slot's columns with the weight planes of its current kt -- the activations of the 3x3 convolution never touch shared
// memory.  Only the two outer kf blocks of D are exchanged with the neighbouring bins through shared memory (32 floats per
// position instead of the 9-tap x 16-channel gather of the FFMA kernel).
//
// Operand format: kind::f16 with fp32-class accuracy through an fp16 hi/lo split, x = hi + lo, hi = fp16(x), lo = fp16(x - hi)
// (11 + 11 mantissa bits, the same as the TF32 split, measured 8.8e-7 abs on K = 48 sums of O(1) values against 5.7e-7 for
// sequential fp32 FMA: profiles/umma_ts_test_r2.txt).  Against TF32 it halves both the TMEM columns of A (two elements per column:
// 8 columns of hi pairs + 8 of lo pairs per 16-channel block) and the number of MMAs (K = 16 per instruction): 9 per step.  TMEM
// is what limits the warpgroups in flight per SM (96 columns each -> 4 or 5 instead of 3), and a CTA's MMAs execute one after the
// other at ~46 cycles each whatever their shape, so both matter.  The weights are scaled by a power of two (exact) so that the lo
// parts stay normal fp16 numbers; the scale is undone in the bias FMA.  fp16 saturates at 65,504: the conversion clamps
// (cvt.rn.satfinite), the point-conv activations it applies to are O(10) for the shipped checkpoint at full-scale input and O(500)
// for the random test weights; GTCRNB_TC=0 selects the FFMA kernels, which have no such limit.
//   TMEM per warpgroup (96 columns): D[0:48], A ring [48:96] (slot s: hi pairs at 48 + 16 s, lo pairs at 56 + 16 s)
//   shared memory: B = W_hi, W_lo fp16 as [kt][2 k-chunks][48 rows = kf*16+co][8 halves] (9 KB); per warpgroup the exchange buffer
//                  X[2][4 quads][100 rows] float4, E2[99][8] (q^2), Gt[3][8] gates; per CTA the 104 TRALite weights
// Point conv 1 / 2 (128 MACs each), PReLUs, TRALite and the channel shuffle stay on the CUDA cores (gt_pc1 / gt_pc2 / gt_gate1, the
// device functions of the FFMA kernel).
// grid (segments, ceil(B / (3 NWG))), 128 NWG threads.
// ================================================================================================
constexpr int GTT_G = 3;
constexpr int GTT_ROWS = GTT_G * kF2;                       // 99 active rows per warpgroup
constexpr int GTT_B_FLOATS = 2 * 3 * 384;                   // [hi/lo][kt][2][48][8 halves] = 9,216 bytes
constexpr int GTT_GW_FLOATS = 104;                          // TRALite weights: TDW_W[3][8], TDW_B[8], TPW_W[8][8], TPW_B[8]
constexpr int GTT_WG_FLOATS = 2 * 4 * 100 * 4 + GTT_ROWS * 8 + GTT_G * 8;   // X, E2, Gt
__host__ __device__ constexpr int gtt_smem_floats(int nwg) { return GTT_B_FLOATS + GTT_GW_FLOATS + nwg * GTT_WG_FLOATS; }
__host__ __device__ constexpr int gtt_min_blocks(int nwg) { return nwg == 1 ? 4 : nwg == 2 ? 2 : 1; }
__host__ __device__ constexpr uint32_t gtt_tmem_cols(int nwg) { return nwg == 1 ? 128 : nwg == 2 ? 256 : 512; }
static_assert(GTT_WG_FLOATS % 4 == 0 && GTT_B_FLOATS % 4 == 0 && GTT_GW_FLOATS % 4 == 0, "16-byte aligned regions");

// STREAM = false: offline, grid (segments, ceil(B / (3 NWG))), every group walks its utterance segment one frame per step.
// STREAM = true:  one 16 ms hop of N streams (StreamGTConvBlock, gtcrn_micro_stream.py:182-262): a persistent grid, every warpgroup
//                 loops over tiles of 3 streams; the two older K blocks come from the stream's point-conv ring in HBM and the TRALite
//                 energy history from its energy ring (gtcrn_stream.cuh state layout) instead of TMEM / registers.  Same operand
//                 split, MMA order and epilogue as the offline kernel, so the decoder blocks of the streaming path are bit-identical
//                 to the offline tensor-core build.
template <int NWG, bool STREAM>
__global__ void __launch_bounds__(128 * NWG, gtt_min_blocks(NWG))
k_gtconv_dec_tc(const float* __restrict__ xin, const float* __restrict__ skip, float* __restrict__ out,
                const __grid_constant__ ParamBlock<gt_chunk<true>()> prm, int T, int seg_len, int B, void* state, int blk, long long N) {
    using O = GtOff<true>;
    static_assert(NWG == 1 || NWG == 2 || (NWG >= 3 && NWG <= 5), "96 TMEM columns per warpgroup, power-of-two allocations");
    constexpr int NTHR = 128 * NWG;
    int slot_old = 0, slot_mid = 1;
    if (STREAM) {
        const gs::Header* hdr = reinterpret_cast<const gs::Header*>(state);
        if (hdr->skip) return;                                        // first wav hop: nothing to enhance yet (whole grid returns)
        slot_old = (int)(hdr->frame_step & 1);                        // ring slot of frame t-2, overwritten with frame t
        slot_mid = slot_old ^ 1;
    }
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) uint64_t bars[NWG];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned wmax_bits;
    float* Bw = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, tig = tid & 127;
    float* Gw = Bw + GTT_B_FLOATS;                                    // TRALite weights
    float* wgm = Gw + GTT_GW_FLOATS + wg * GTT_WG_FLOATS;
    float4* X = reinterpret_cast<float4*>(wgm);                       // [2][4][100]
    float* E2 = wgm + 2 * 4 * 100 * 4;                                // [99][8]
    float* Gt = E2 + GTT_ROWS * 8;                                    // [3][8]

    // ---- prologue: power-of-two weight scale, fp16 hi/lo weight planes, TRALite weights, tensor memory
    if (tid == 0) wmax_bits = 0u;
    __syncthreads();
    {
        unsigned mx = 0u;
        for (int i = tid; i < 2304; i += NTHR) mx = max(mx, __float_as_uint(fabsf(prm.w[O::MID_W + i])));
        atomicMax(&wmax_bits, mx);
    }
    __syncthreads();
    // |w| < 2^(e+1)  =>  |w| * 2^(13-e) < 2^14: hi parts well inside fp16, lo parts (~2^-11 of that) normal down to |w| ~ 2^-13 max|w|
    const int wexp = max(14, min(253, (int)(wmax_bits >> 23)));
    const float wscale = __uint_as_float((unsigned)(127 + 13 + 127 - wexp) << 23);
    const float winv = __uint_as_float((unsigned)(wexp - 13) << 23);
    {
        __half* Bh = reinterpret_cast<__half*>(Bw);
        for (int i = tid; i < 2304; i += NTHR) {
            const int kt = i / 768, r = i - kt * 768, kf = r >> 8, ci = (r >> 4) & 15, co = r & 15;
            const float w = prm.w[O::MID_W + ((kt * 3 + kf) * 16 + ci) * 16 + co] * wscale;
            const __half hi = __float2half_rn(w);
            const int o = kt * 768 + (ci >> 3) * 384 + (kf * 16 + co) * 8 + (ci & 7);      // in halves
            Bh[o] = hi;
            Bh[2304 + o] = __float2half_rn(w - __half2float(hi));
        }
    }
    for (int i = tid; i < GTT_GW_FLOATS; i += NTHR)
        Gw[i] = i < 24 ? prm.w[O::TDW_W + i] : i < 32 ? prm.w[O::TDW_B + i - 24] : i < 96 ? prm.w[O::TPW_W + i - 32] : prm.w[O::TPW_B + i - 96];
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, gtt_tmem_cols(NWG));
    if (tid < NWG) tc::mbar_init(&bars[tid], 1);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tm = tmem_base_s + 96 * wg, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    if (!STREAM) {
        // zero history: frames before the start of the utterance (or of the warm-up) contribute nothing
        uint32_t z[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) z[c] = 0u;
#pragma unroll
        for (int j = 0; j < 6; ++j) tc::tmem_st8u(tm + lane_base + 48 + 8 * j, z);
        tc::tmem_wait_st();
    }
    const uint64_t bdesc0 = tc::make_desc(tc::smem_u32(Bw), 768, 128);       // LBO = one k-chunk plane (48 rows x 16 B), SBO = 8 rows
    // Who does what between the two barriers of a step (clock64 trace, tools/gpu_dec_trace.sh): issuing 9 MMAs + commit keeps one
    // thread busy for ~1,000 cycles, the TRALite gate of the previous frame is a ~1,000-cycle dependent chain on 24 threads, and the
    // MMAs complete ~50 cycles after the last one is issued.  Originally thread 0 did both one after the other (it is row (0, 0), one
    // of the gate threads: 6,700 cycles per step).  Now thread 0 issues and the gate runs on lanes 0..23 of WARP 3, whose rows 99..127
    // are idle: the two overlap, and warps 0-2 go straight to the completion wait.
    const bool issuer = tig == 0;
    const int bar_id = 1 + wg;
    uint64_t* bar = &bars[wg];
    uint32_t parity = 0;

    const bool active = tig < GTT_ROWS;
    const int rr = active ? tig : GTT_ROWS - 1;                          // idle lanes shadow the last row (no stores, no smem writes)
    const int g = rr / kF2, f = rr - g * kF2;
    const bool is_gate_thread = tig >= 96 && tig < 96 + 8 * GTT_G;       // warp 3, lanes 0..23: gate thread (gg, gc)
    const int gg = is_gate_thread ? (tig - 96) >> 3 : 0, gc = tig & 7;
    const unsigned gmask = 0xFFu << (8 * gg);
    constexpr int FS4 = kF2 * 4;                                         // float4 per frame

    // fp16 hi/lo split of a 16-channel row into ring slot `slot` of the A operand (8 columns of hi pairs, 8 of lo pairs)
    auto store_slot = [&](int slot, const float (&p)[16]) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            hi[c] = tc::pack_f16x2(p[2 * c], p[2 * c + 1]);
            lo[c] = tc::pack_f16x2(p[2 * c] - tc::f16lo_to_f32(hi[c]), p[2 * c + 1] - tc::f16hi_to_f32(hi[c]));
        }
        tc::tmem_st8u(tm + lane_base + 48 + 16 * slot, hi);
        tc::tmem_st8u(tm + lane_base + 56 + 16 * slot, lo);
    };
    // the 9 MMAs of one frame; r = ring slot of the newest frame: kt = 0 <-> frame t-2 <-> slot r+1, kt = 2 <-> frame t <-> slot r
    auto issue_mmas = [&](int r) {
        tc::fence_after();
        constexpr uint32_t id48 = tc::make_idesc_f16(128, 48);
        uint32_t acc = 0;
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {                              // small terms first
            int s = r + 1 + kt;
            s -= s >= 3 ? 3 : 0;
            const uint64_t bh = bdesc0 + (uint64_t)(kt * 96);         // 1,536 bytes per kt >> 4
            tc::mma_f16_ts(tm, tm + 56 + 16 * s, bh, id48, acc); acc = 1;             // lo * W_hi
            tc::mma_f16_ts(tm, tm + 48 + 16 * s, bh + 288, id48, 1);                  // hi * W_lo (4,608 bytes further)
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
            int s = r + 1 + kt;
            s -= s >= 3 ? 3 : 0;
            tc::mma_f16_ts(tm, tm + 48 + 16 * s, bdesc0 + (uint64_t)(kt * 96), id48, 1);   // hi * W_hi
        }
        tc::commit(bar);
    };
    // accumulator row back (kf = 0 / kf = 2 blocks to the neighbouring bins through shared memory), first half: up to the barrier
    auto fetch_and_exchange = [&](float (&m)[16]) {
        {
            float d0[16], d2[16];
            tc::tmem_ld16_nowait(tm + lane_base, d0);
            tc::tmem_ld16_nowait(tm + lane_base + 32, d2);
            tc::tmem_ld16_nowait(tm + lane_base + 16, m);                 // all three blocks in one tensor-memory round trip
            tc::tmem_wait_ld(d0);
            tc::tmem_wait_ld(d2);
            tc::tmem_wait_ld(m);
            if (active) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    X[qd * 100 + rr] = make_float4(d0[4 * qd], d0[4 * qd + 1], d0[4 * qd + 2], d0[4 * qd + 3]);
                    X[400 + qd * 100 + rr] = make_float4(d2[4 * qd], d2[4 * qd + 1], d2[4 * qd + 2], d2[4 * qd + 3]);
                }
            }
        }
        tc::fence_before();                                           // orders the tcgen05.ld above before the next step's MMAs
    };
    // second half: neighbours' blocks, weight scale + bias, PReLU, point conv 2, q^2 for the TRALite energy
    auto epilogue = [&](float (&m)[16], float (&q)[8]) {
        if (f > 0) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float4 u = X[qd * 100 + rr - 1];
                add_pair(m[4 * qd], m[4 * qd + 1], u.x, u.y); add_pair(m[4 * qd + 2], m[4 * qd + 3], u.z, u.w);
            }
        }
        if (f < kF2 - 1) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const float4 u = X[400 + qd * 100 + rr + 1];
                add_pair(m[4 * qd], m[4 * qd + 1], u.x, u.y); add_pair(m[4 * qd + 2], m[4 * qd + 3], u.z, u.w);
            }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) m[c] = fmaf(m[c], winv, prm.w[O::MID_B + c]);     // undo the weight scale, add the bias
        gt_pc2<true>(prm, m, q);
        if (active) {
            float4* e4 = reinterpret_cast<float4*>(E2 + rr * 8);
            e4[0] = make_float4(q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]);
            e4[1] = make_float4(q[4] * q[4], q[5] * q[5], q[6] * q[6], q[7] * q[7]);
        }
    };
    // TRALite energy + gate of one finished frame (its q^2 are in E2): thread (g, c < 8) owns channel c of group g -- its energy
    // history (ea, eb = frames -2, -1) is in registers, the depthwise temporal conv is one value per thread and the 8 x 8
    // pointwise conv exchanges them with shuffles among the group's 8 lanes.  Four interleaved partial sums in a fixed order:
    // deterministic, independent of batch composition and segmentation.  Returns with eb = the new energy.
    float ea = 0.f, eb = 0.f;
    auto gate_phase = [&]() {
        if (is_gate_thread) {
            const int c = gc;
            const float* e2c = E2 + gg * kF2 * 8 + c;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int ff = 0; ff < 32; ff += 4) {
                s0 += e2c[ff * 8]; s1 += e2c[(ff + 1) * 8]; s2 += e2c[(ff + 2) * 8]; s3 += e2c[(ff + 3) * 8];
            }
            s0 += e2c[32 * 8];
            const float en = ((s0 + s1) + (s2 + s3)) / 33.0f;
            float y = Gw[24 + c];
            y = fmaf(Gw[c], ea, y);
            y = fmaf(Gw[8 + c], eb, y);
            y = fmaf(Gw[16 + c], en, y);
            ea = eb; eb = en;
            float v = Gw[96 + c];
            const int lane0 = 8 * gg;
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) v = fmaf(Gw[32 + ci * 8 + c], __shfl_sync(gmask, y, lane0 + ci), v);
            Gt[gg * 8 + c] = 1.0f / (1.0f + expf(-v));
        }
    };
    // gate, channel shuffle [q0 g0, x2_0, q1 g1, x2_1, ...], store
    auto store_row = [&](float4* o4, const float (&q)[8], const float4& x2a, const float4& x2b) {
        const float4 g0 = reinterpret_cast<const float4*>(Gt + g * 8)[0];
        const float4 g1 = reinterpret_cast<const float4*>(Gt + g * 8)[1];
        o4[0] = make_float4(q[0] * g0.x, x2a.x, q[1] * g0.y, x2a.y);
        o4[1] = make_float4(q[2] * g0.z, x2a.z, q[3] * g0.w, x2a.w);
        o4[2] = make_float4(q[4] * g1.x, x2b.x, q[5] * g1.y, x2b.y);
        o4[3] = make_float4(q[6] * g1.z, x2b.z, q[7] * g1.w, x2b.w);
    };

    if (STREAM) {
        const long long ntiles = (N + GTT_G - 1) / GTT_G;
        for (long long tile = (long long)blockIdx.x * NWG + wg; tile < ntiles; tile += (long long)gridDim.x * NWG) {
            const long long n_raw = tile * GTT_G + g;
            const bool valid = active && n_raw < N;
            const long long n = min(n_raw, N - 1);
            float* sb = gs::stream_base(state, n);
            float* ring = sb + gs::kGtP + blk * (2 * kF2 * 16) + f * 16;
            {
                // the operands of this warpgroup's next tile start their trip from HBM now (a tile is ~7 us of dependent work)
                const long long nn = min((tile + (long long)gridDim.x * NWG) * GTT_G + g, N - 1);
                const float* rn = gs::stream_base(state, nn) + gs::kGtP + blk * (2 * kF2 * 16) + f * 16;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(xin + (nn * kF2 + f) * 16));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(skip + (nn * kF2 + f) * 16));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rn));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(rn + kF2 * 16));
            }
            float4 cx[4];
            {
                const float4* px = reinterpret_cast<const float4*>(xin + (n * kF2 + f) * 16);
                const float4* ps = reinterpret_cast<const float4*>(skip + (n * kF2 + f) * 16);
                const float4* h2 = reinterpret_cast<const float4*>(ring + slot_old * (kF2 * 16));
                const float4* h1 = reinterpret_cast<const float4*>(ring + slot_mid * (kF2 * 16));
                float4 a[4], c[4], u2[4], u1[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) { a[v] = px[v]; c[v] = ps[v]; u2[v] = h2[v]; u1[v] = h1[v]; }
                if (is_gate_thread) {                                         // energy history of stream (tile, gg), channel gc
                    const float* sg = gs::stream_base(state, min(tile * GTT_G + gg, N - 1));
                    ea = sg[gs::kTraE + blk * 16 + slot_old * 8 + gc];
                    eb = sg[gs::kTraE + blk * 16 + slot_mid * 8 + gc];
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) cx[v] = make_float4(a[v].x + c[v].x, a[v].y + c[v].y, a[v].z + c[v].z, a[v].w + c[v].w);
                {
                    const float p2[16] = {u2[0].x, u2[0].y, u2[0].z, u2[0].w, u2[1].x, u2[1].y, u2[1].z, u2[1].w,
                                          u2[2].x, u2[2].y, u2[2].z, u2[2].w, u2[3].x, u2[3].y, u2[3].z, u2[3].w};
                    store_slot(0, p2);
                }
                {
                    const float p1[16] = {u1[0].x, u1[0].y, u1[0].z, u1[0].w, u1[1].x, u1[1].y, u1[1].z, u1[1].w,
                                          u1[2].x, u1[2].y, u1[2].z, u1[2].w, u1[3].x, u1[3].y, u1[3].z, u1[3].w};
                    store_slot(1, p1);
                }
            }
            {
                const float x1[8] = {cx[0].x, cx[0].y, cx[0].z, cx[0].w, cx[1].x, cx[1].y, cx[1].z, cx[1].w};
                float p[16];
                gt_pc1<true>(prm, x1, p);
                store_slot(2, p);
                if (valid) {                                              // the new frame replaces frame t-2 in the stream's ring
                    float4* newest = reinterpret_cast<float4*>(ring + slot_old * (kF2 * 16));
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) newest[qd] = make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]);
                }
            }
            tc::tmem_wait_st();
            tc::fence_before();
            tc::bar_sync(bar_id, 128);
            if (issuer) issue_mmas(2);                                    // fixed slots: kt = slot
            tc::mbar_wait(bar, parity);
            parity ^= 1;
            tc::fence_after();
            float m[16], q[8];
            fetch_and_exchange(m);
            tc::bar_sync(bar_id, 128);
            epilogue(m, q);
            tc::bar_sync(bar_id, 128);
            gate_phase();
            if (is_gate_thread && tile * GTT_G + gg < N) gs::stream_base(state, tile * GTT_G + gg)[gs::kTraE + blk * 16 + slot_old * 8 + gc] = eb;
            tc::bar_sync(bar_id, 128);
            if (valid) store_row(reinterpret_cast<float4*>(out + (n * kF2 + f) * 16), q, cx[2], cx[3]);
        }
    } else {
        const int b_raw = blockIdx.y * (GTT_G * NWG) + wg * GTT_G + g;
        const bool store_ok = active && b_raw < B;
        const int b = min(b_raw, B - 1);
        const int s0 = blockIdx.x * seg_len, s1 = min(T, s0 + seg_len);
        if (s0 < s1) {
            const int tw0 = max(0, s0 - GT_WARMUP);
            const int64_t ubase = (int64_t)b * T * kF2 * 16 + f * 16;
            const float4* px = reinterpret_cast<const float4*>(xin + ubase);
            const float4* ps = reinterpret_cast<const float4*>(skip + ubase);
            float4* po = reinterpret_cast<float4*>(out + ubase);
            // The loop is software-pipelined by one frame: the TRALite gate of frame t-1 is evaluated by its 24 threads while the
            // MMAs of frame t run, and frame t-1 is stored after the barrier that publishes the gate -- two warpgroup barriers per frame.
            int r = tw0 % 3;                                              // ring slot of the current frame
            float qp[8];                                                  // previous frame: point-conv-2 output and pass-through half
#pragma unroll
            for (int c = 0; c < 8; ++c) qp[c] = 0.f;
#ifdef GTCRNB_DEC_TRACE
            long long* trc = reinterpret_cast<long long*>(state);
            const bool tracer = trc && blockIdx.x == 0 && blockIdx.y == 0 && (tig == 0 || tig == 40) ;
            const int trow = (tid >> 7) * 2 + (tig == 40);
#define TRC(i) do { if (tracer && t >= tw0 + 16 && t < tw0 + 48) { long long c_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c_)); trc[((trow * 32 + (t - tw0 - 16)) * 12) + (i)] = c_; } } while (0)
#else
#define TRC(i) do { } while (0)
#endif
            // The point-conv half (channels 0..7) of x and skip of the NEXT frame is requested right after the stores of step t, so that
            // its L2 latency (~1,200 cycles on the trace) hides behind the epilogue instead of opening the next step; the pass-through
            // half of a frame is only needed for its store one step later and is requested during that step's MMA wait.  The frame after next is on
            // its way to L2.
            float4 a[2], c[2];
#pragma unroll
            for (int v = 0; v < 2; ++v) { a[v] = px[(int64_t)tw0 * FS4 + v]; c[v] = ps[(int64_t)tw0 * FS4 + v]; }
            if (tw0 + 1 < s1) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(px + (int64_t)(tw0 + 1) * FS4));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + (int64_t)(tw0 + 1) * FS4));
            }
            for (int t = tw0; t < s1; ++t) {
                TRC(0);
                float4 cx[2];                                             // x + skip of this frame, point-conv half
                const int tp = t > tw0 ? t - 1 : t;
                const float4 a2 = px[(int64_t)tp * FS4 + 2], a3 = px[(int64_t)tp * FS4 + 3], c2 = ps[(int64_t)tp * FS4 + 2], c3 = ps[(int64_t)tp * FS4 + 3];
                {
                    if (t + 2 < s1) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(px + (int64_t)(t + 2) * FS4));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + (int64_t)(t + 2) * FS4));
                    }
#pragma unroll
                    for (int v = 0; v < 2; ++v) cx[v] = make_float4(a[v].x + c[v].x, a[v].y + c[v].y, a[v].z + c[v].z, a[v].w + c[v].w);
                }
                TRC(1);
                // ---- point conv 1 on the first channel half, fp16 hi/lo split, into ring slot r of the A operand
                {
                    const float x1[8] = {cx[0].x, cx[0].y, cx[0].z, cx[0].w, cx[1].x, cx[1].y, cx[1].z, cx[1].w};
                    float p[16];
                    gt_pc1<true>(prm, x1, p);
                    store_slot(r, p);
                }
                TRC(2);
                tc::tmem_wait_st();
                tc::fence_before();
                TRC(3);
                tc::bar_sync(bar_id, 128);                                // A(t) complete; also: E2(t-1) visible to the gate threads
                TRC(4);
                if (issuer) issue_mmas(r);
                TRC(5);
                // pass-through half (channels 8..15) of frame t-1, stored below: requested here, where few registers are live and the
                // MMAs / the gate give its L2 latency ~1,000 cycles (it shares its 128-byte line with the half loaded a step ago)
                if (t > tw0) gate_phase();                                // frame t-1, while the tensor core works on frame t
                TRC(6);
                tc::mbar_wait(bar, parity);
                parity ^= 1;
                tc::fence_after();
                TRC(7);
                float m[16];
                fetch_and_exchange(m);
                TRC(8);
                tc::bar_sync(bar_id, 128);                                // exchange buffer and the gate of frame t-1 are visible
                TRC(9);
                if (t > tw0 && store_ok && t - 1 >= s0)
                    store_row(po + (int64_t)(t - 1) * FS4, qp, make_float4(a2.x + c2.x, a2.y + c2.y, a2.z + c2.z, a2.w + c2.w),
                              make_float4(a3.x + c3.x, a3.y + c3.y, a3.z + c3.z, a3.w + c3.w));
                TRC(10);
                if (t + 1 < s1) {
#pragma unroll
                    for (int v = 0; v < 2; ++v) { a[v] = px[(int64_t)(t + 1) * FS4 + v]; c[v] = ps[(int64_t)(t + 1) * FS4 + v]; }
                }
                epilogue(m, qp);
                TRC(11);
                r = r == 2 ? 0 : r + 1;
            }
            // ---- drain: gate and store of the last frame
            const float4 a2 = px[(int64_t)(s1 - 1) * FS4 + 2], a3 = px[(int64_t)(s1 - 1) * FS4 + 3], c2 = ps[(int64_t)(s1 - 1) * FS4 + 2], c3 = ps[(int64_t)(s1 - 1) * FS4 + 3];
            tc::bar_sync(bar_id, 128);
            gate_phase();
            tc::bar_sync(bar_id, 128);
            if (store_ok)
                store_row(po + (int64_t)(s1 - 1) * FS4, qp, make_float4(a2.x + c2.x, a2.y + c2.y, a2.z + c2.z, a2.w + c2.w),
                          make_float4(a3.x + c3.x, a3.y + c3.y, a3.z + c3.z, a3.w + c3.w));
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, gtt_tmem_cols(NWG));
}

// one decoder block of a streaming hop on the tensor cores: persistent grid of one CTA per SM (or fewer for small N)
inline cudaError_t launch_dec_tc_stream(const float* x, const float* skip, float* out, const ParamBlock<gt_chunk<true>()>& prm, void* state, int blk,
                                        int64_t N, int sm_count, cudaStream_t st) {
    constexpr int NWG = 4;
    const long long ntiles = (N + GTT_G - 1) / GTT_G;
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((ntiles + NWG - 1) / NWG, sm_count));
    int T = 1, seg = 1, B = 0; long long n = N;
    void* args[] = {(void*)&x, (void*)&skip, (void*)&out, (void*)&prm, (void*)&T, (void*)&seg, (void*)&B, (void*)&state, (void*)&blk, (void*)&n};
    return cudaLaunchKernel((const void*)k_gtconv_dec_tc<NWG, true>, dim3(grid), dim3(128 * NWG), args, gtt_smem_floats(NWG) * sizeof(float), st);
}

}  // namespace gk
