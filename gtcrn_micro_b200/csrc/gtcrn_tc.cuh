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
#include "gtcrn_kernels.cuh"
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
__host__ __device__ constexpr int tct_min_blocks(int nwg) { return nwg == 1 ? 7 : nwg == 2 ? 4 : 2; }

// one pointwise 16 -> 16 convolution of the warpgroup's 128 rows, first half: v (registers) -> A in TMEM, MMAs, completion wait.
// STACK: 4 MMAs, the accumulator holds hi*w_hi + lo*w_hi in columns 0..15 and hi*w_lo in 16..31 (dependent chain of 4);
// otherwise 6 MMAs accumulate all three terms into columns 0..15 (chain of 6, half the tcgen05.ld traffic).
template <bool STACK>
__device__ __forceinline__ void tct_conv16_issue(const float (&v)[16], uint32_t tm, uint32_t lane_base, uint64_t bdesc, uint64_t* bar,
                                                  uint32_t& parity, int bar_id, bool issuer) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {                                  // in halves: 16 instead of 32 staging registers
        float hi[8], lo[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { hi[c] = tc::tf32_hi(v[8 * h + c]); lo[c] = v[8 * h + c] - hi[c]; }
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

        for (int t0 = tw0; t0 < s1; t0 += 32) {
            const int t = t0 + lane;
            float x[16];
            if (t < s1) {
                const float4* px = reinterpret_cast<const float4*>(xcol + (int64_t)t * (kF2 * 16));
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 u = px[v];
                    x[4 * v] = u.x; x[4 * v + 1] = u.y; x[4 * v + 2] = u.z; x[4 * v + 3] = u.w;
                }
                if (t + 32 < s1) asm volatile("prefetch.global.L2 [%0];" ::"l"(xcol + (int64_t)(t + 32) * (kF2 * 16)));
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) x[c] = 0.f;
            }
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int d = 1 << k, coff = (2 << k) - 2, wo = k * blob::TCN_SIZE;
                float y[16];
                tct_conv16_issue<STACK>(x, tm, lane_base, bdesc0 + (uint64_t)(2 * k) * 128, bar, parity, 1 + wg, issuer);
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
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) Y[qd * 32 + lane] = make_float4(y[4 * qd], y[4 * qd + 1], y[4 * qd + 2], y[4 * qd + 3]);
                __syncwarp();
                {
                    // depthwise dilated taps, one channel quad at a time (8 instead of 32 tap registers live)
                    const float a2 = prm.w[wo + blob::TCN_SLOPE + 1];
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        const float4 ta = lane >= 2 * d ? Y[qd * 32 + lane - 2 * d] : H[qd * 30 + coff + lane];
                        const float4 tb = lane >= d ? Y[qd * 32 + lane - d] : H[qd * 30 + coff + lane + d];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {                    // scalar on purpose (see tcn_dw)
                            const int c = 4 * qd + i;
                            float v = prm.w[wo + blob::TCN_B2 + c];
                            v = fmaf(prm.w[wo + blob::TCN_W2 + c], f4get(ta, i), v);
                            v = fmaf(prm.w[wo + blob::TCN_W2 + 16 + c], f4get(tb, i), v);
                            v = fmaf(prm.w[wo + blob::TCN_W2 + 32 + c], y[c], v);
                            y[c] = prelu(v, a2);
                        }
                    }
                }
                tct_conv16_issue<STACK>(y, tm, lane_base, bdesc0 + (uint64_t)(2 * k + 1) * 128, bar, parity, 1 + wg, issuer);
                // history <- the last 2d conv1 rows of this tile.  The warpgroup barrier inside the call above separates these writes
                // from every lane's tap reads, the one of the next conv1 separates the Y reads from the next Y writes.
                if (lane < 8 * d) {
                    const int qd = lane / (2 * d), j = lane - qd * 2 * d;
                    H[qd * 30 + coff + j] = Y[qd * 32 + 32 - 2 * d + j];
                }
                if (lane + 32 < 8 * d) {                                    // d = 8: 64 float4 to move
                    const int i = lane + 32, qd = i / (2 * d), j = i - qd * 2 * d;
                    H[qd * 30 + coff + j] = Y[qd * 32 + 32 - 2 * d + j];
                }
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
            }
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
// its last three steps: they sit in a 3-slot ring of TMEM columns (A_hi / A_lo, 16 columns per slot), every step overwrites the
// oldest slot with the new frame and pairs each slot's columns with the weight planes of its current kt -- the activations of the
// 3x3 convolution never touch shared memory.  Only the two outer kf blocks of D are exchanged with the neighbouring bins through
// shared memory (32 floats per position instead of the 9-tap x 16-channel gather of the FFMA kernel).
//   TMEM per warpgroup (144 columns): D[0:48], A_hi[48:96] (slot s at 48 + 16 s), A_lo[96:144]; one CTA of 3 warpgroups per SM
//   shared memory: B = W_hi, W_lo as [kt][4 k-planes][48 rows = kf*16+co][4] (18 KB); per warpgroup the exchange buffer
//                  X[2][4 quads][100 rows] float4, E2[99][8] (q^2), Ev[3 groups][3 slots][8] TRA energies, Gt[3][8] gates
// Point conv 1 / 2 (128 MACs each), PReLUs, TRALite and the channel shuffle stay on the CUDA cores (gt_pc1 / gt_pc2 / gt_gate1, the
// device functions of the FFMA kernel).
// grid (segments, ceil(B / 9)), 384 threads.
// ================================================================================================
constexpr int GTT_NWG = 3;
constexpr int GTT_G = 3;
constexpr int GTT_ROWS = GTT_G * kF2;                       // 99 active rows per warpgroup
constexpr int GTT_THREADS = 128 * GTT_NWG;
constexpr int GTT_B_FLOATS = 2 * 3 * 768;                   // [hi/lo][kt][4][48][4]
constexpr int GTT_WG_FLOATS = 2 * 4 * 100 * 4 + GTT_ROWS * 8 + GTT_G * 3 * 8 + GTT_G * 8 + 8;   // X, E2, Ev, Gt (+ pad to 16 B)
__host__ __device__ constexpr int gtt_smem_floats() { return GTT_B_FLOATS + GTT_NWG * GTT_WG_FLOATS; }
static_assert(GTT_WG_FLOATS % 4 == 0 && GTT_B_FLOATS % 4 == 0, "16-byte aligned regions");

__global__ void __launch_bounds__(GTT_THREADS, 1)
k_gtconv_dec_tc(const float* __restrict__ xin, const float* __restrict__ skip, float* __restrict__ out,
                const __grid_constant__ ParamBlock<gt_chunk<true>()> prm, int T, int seg_len, int B) {
    using O = GtOff<true>;
    extern __shared__ float4 smem4[];
    __shared__ __align__(8) uint64_t bars[GTT_NWG];
    __shared__ uint32_t tmem_base_s;
    float* Bw = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, tig = tid & 127;
    float* wgm = Bw + GTT_B_FLOATS + wg * GTT_WG_FLOATS;
    float4* X = reinterpret_cast<float4*>(wgm);                       // [2][4][100]
    float* E2 = wgm + 2 * 4 * 100 * 4;                                // [99][8]
    float* Ev = E2 + GTT_ROWS * 8;                                    // [3][3][8]
    float* Gt = Ev + GTT_G * 3 * 8;                                   // [3][8]

    // ---- prologue: weight planes (3xTF32 split, round to nearest), zeroed TRA history, tensor memory
    for (int i = tid; i < 2304; i += GTT_THREADS) {
        const int kt = i / 768, r = i - kt * 768, kf = r >> 8, ci = (r >> 4) & 15, co = r & 15;
        const float w = prm.w[O::MID_W + ((kt * 3 + kf) * 16 + ci) * 16 + co];
        float hi, lo;
        tf32_split_rn(w, hi, lo);
        const int o = kt * 768 + (ci >> 2) * 192 + (kf * 16 + co) * 4 + (ci & 3);
        Bw[o] = hi;
        Bw[2304 + o] = lo;
    }
    for (int i = tig; i < GTT_G * 3 * 8; i += 128) Ev[i] = 0.f;
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    if (tid < GTT_NWG) tc::mbar_init(&bars[tid], 1);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tm = tmem_base_s + 144 * wg, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    {
        // zero history: frames before the start of the utterance (or of the warm-up) contribute nothing
        float z[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) z[c] = 0.f;
#pragma unroll
        for (int j = 0; j < 12; ++j) tc::tmem_st8(tm + lane_base + 48 + 8 * j, z);
        tc::tmem_wait_st();
    }
    const uint64_t bdesc0 = tc::make_desc(tc::smem_u32(Bw), 768, 128);       // LBO = one k-plane (48 rows x 16 B), SBO = 8 rows
    const bool issuer = tig == 0;
    const int bar_id = 1 + wg;
    uint64_t* bar = &bars[wg];
    uint32_t parity = 0;

    const bool active = tig < GTT_ROWS;
    const int rr = active ? tig : GTT_ROWS - 1;                          // idle lanes shadow the last row (no stores, no smem writes)
    const int g = rr / kF2, f = rr - g * kF2;
    const int b_raw = blockIdx.y * (GTT_G * GTT_NWG) + wg * GTT_G + g;
    const bool store_ok = active && b_raw < B;
    const int b = min(b_raw, B - 1);
    const int s0 = blockIdx.x * seg_len, s1 = min(T, s0 + seg_len);
    if (s0 < s1) {
        const int tw0 = max(0, s0 - GT_WARMUP);
        const int64_t ubase = (int64_t)b * T * kF2 * 16 + f * 16;
        const float4* px = reinterpret_cast<const float4*>(xin + ubase);
        const float4* ps = reinterpret_cast<const float4*>(skip + ubase);
        float4* po = reinterpret_cast<float4*>(out + ubase);
        constexpr int FS4 = kF2 * 4;                                     // float4 per frame
        const bool is_gate_thread = active && f < 8;
        const unsigned gmask = 0xFFu << ((g * kF2) & 31);                // lanes of (g, c = 0..7): 33 g + c = g + c (mod 32)

        float4 nx[8];                                                     // next frame: x quads 0..3, skip quads 0..3
#pragma unroll
        for (int v = 0; v < 4; ++v) { nx[v] = px[(int64_t)tw0 * FS4 + v]; nx[4 + v] = ps[(int64_t)tw0 * FS4 + v]; }
        int r = tw0 % 3;                                                  // ring slot of the current frame
        for (int t = tw0; t < s1; ++t) {
            float4 cx[4];
#pragma unroll
            for (int v = 0; v < 4; ++v)
                cx[v] = make_float4(nx[v].x + nx[4 + v].x, nx[v].y + nx[4 + v].y, nx[v].z + nx[4 + v].z, nx[v].w + nx[4 + v].w);
            if (t + 1 < s1) {
#pragma unroll
                for (int v = 0; v < 4; ++v) { nx[v] = px[(int64_t)(t + 1) * FS4 + v]; nx[4 + v] = ps[(int64_t)(t + 1) * FS4 + v]; }
            }
            // ---- point conv 1 on the first channel half, split, into ring slot r of the A operand
            {
                const float x1[8] = {cx[0].x, cx[0].y, cx[0].z, cx[0].w, cx[1].x, cx[1].y, cx[1].z, cx[1].w};
                float p[16];
                gt_pc1<true>(prm, x1, p);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float hi[8], lo[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) { hi[c] = tc::tf32_hi(p[8 * h + c]); lo[c] = p[8 * h + c] - hi[c]; }
                    tc::tmem_st8(tm + lane_base + 48 + 16 * r + 8 * h, hi);
                    tc::tmem_st8(tm + lane_base + 96 + 16 * r + 8 * h, lo);
                }
            }
            tc::tmem_wait_st();
            tc::fence_before();
            tc::bar_sync(bar_id, 128);
            if (issuer) {
                tc::fence_after();
                constexpr uint32_t id48 = tc::make_idesc_tf32(128, 48);
                uint32_t acc = 0;
#pragma unroll
                for (int kt = 0; kt < 3; ++kt) {
                    int s = r + 1 + kt;                                   // kt = 0 <-> frame t-2 <-> slot r+1, kt = 2 <-> frame t <-> slot r
                    s -= s >= 3 ? 3 : 0;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint64_t bh = bdesc0 + (uint64_t)((kt * 768 + ks * 384) >> 2);          // (floats * 4 B) >> 4
                        const uint64_t bl = bh + (uint64_t)(2304 >> 2);
                        tc::mma_tf32_ts(tm, tm + 96 + 16 * s + 8 * ks, bh, id48, acc); acc = 1;       // small terms first: lo * W_hi
                        tc::mma_tf32_ts(tm, tm + 48 + 16 * s + 8 * ks, bl, id48, 1);                  // hi * W_lo
                    }
                }
#pragma unroll
                for (int kt = 0; kt < 3; ++kt) {
                    int s = r + 1 + kt;
                    s -= s >= 3 ? 3 : 0;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        tc::mma_tf32_ts(tm, tm + 48 + 16 * s + 8 * ks, bdesc0 + (uint64_t)((kt * 768 + ks * 384) >> 2), id48, 1);   // hi * W_hi
                }
                tc::commit(bar);
            }
            tc::mbar_wait(bar, parity);
            parity ^= 1;
            tc::fence_after();
            // ---- accumulator row back: kf = 0 and kf = 2 blocks go to the neighbouring bins through shared memory
            float m[16];
            {
                float d0[16], d2[16];
                tc::tmem_ld16_nowait(tm + lane_base, d0);
                tc::tmem_ld16_nowait(tm + lane_base + 16, m);
                tc::tmem_ld16_nowait(tm + lane_base + 32, d2);
                tc::tmem_wait_ld(d0);
                tc::tmem_wait_ld(m);
                tc::tmem_wait_ld(d2);
                if (active) {
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) {
                        X[qd * 100 + rr] = make_float4(d0[4 * qd], d0[4 * qd + 1], d0[4 * qd + 2], d0[4 * qd + 3]);
                        X[400 + qd * 100 + rr] = make_float4(d2[4 * qd], d2[4 * qd + 1], d2[4 * qd + 2], d2[4 * qd + 3]);
                    }
                }
            }
            tc::fence_before();                                           // orders the tcgen05.ld above before the next step's MMAs
            tc::bar_sync(bar_id, 128);
#pragma unroll
            for (int c = 0; c < 16; c += 2) add_pair(m[c], m[c + 1], prm.w[O::MID_B + c], prm.w[O::MID_B + c + 1]);
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
            // ---- PReLU, point conv 2, TRALite energy and gate
            float q[8];
            gt_pc2<true>(prm, m, q);
            if (active) {
                float4* e4 = reinterpret_cast<float4*>(E2 + rr * 8);
                e4[0] = make_float4(q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]);
                e4[1] = make_float4(q[4] * q[4], q[5] * q[5], q[6] * q[6], q[7] * q[7]);
            }
            tc::bar_sync(bar_id, 128);
            if (is_gate_thread) {
                const int c = f;
                int r1 = r - 1; r1 += r1 < 0 ? 3 : 0;
                int r2 = r1 - 1; r2 += r2 < 0 ? 3 : 0;
                float* ev = Ev + g * 24;
                ev[r * 8 + c] = gt_energy(E2 + g * kF2 * 8 + c);
                __syncwarp(gmask);
                Gt[g * 8 + c] = gt_gate1<true>(prm, ev + r2 * 8, ev + r1 * 8, ev + r * 8, c);
            }
            tc::bar_sync(bar_id, 128);
            // ---- gate, channel shuffle [q0 g0, x2_0, q1 g1, x2_1, ...], store
            if (store_ok && t >= s0) {
                const float4 g0 = reinterpret_cast<const float4*>(Gt + g * 8)[0];
                const float4 g1 = reinterpret_cast<const float4*>(Gt + g * 8)[1];
                float4* o4 = po + (int64_t)t * FS4;
                o4[0] = make_float4(q[0] * g0.x, cx[2].x, q[1] * g0.y, cx[2].y);
                o4[1] = make_float4(q[2] * g0.z, cx[2].z, q[3] * g0.w, cx[2].w);
                o4[2] = make_float4(q[4] * g1.x, cx[3].x, q[5] * g1.y, cx[3].y);
                o4[3] = make_float4(q[6] * g1.z, cx[3].z, q[7] * g1.w, cx[3].w);
            }
            r = r == 2 ? 0 : r + 1;
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace gk
