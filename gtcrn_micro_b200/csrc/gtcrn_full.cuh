// Kernels of the ORIGINAL-GTCRN-style variant (BASELINE configs[3]: unfold SFE, grouped convs, dilated depthwise GTConv,
// GRU-based TRA, DPGRNN).  PARITY UNPINNED: /root/reference contains no implementation of these blocks -- only the commented-out
// lines gtcrn_micro/models/gtcrn_micro.py:187,199,346,361,381,391,417-418,429,448,494 and the prose at :5 ("Replacing DPGRNN
// approach with grouped TCNs") -- so the checker is the builder-written oracle/gtcrn_full_oracle.py.
//
// The per-frame ends of the path (STFT, ERB, en0 / en1, de3 / de4, mask, iSTFT) are the VAR = 1 instantiations of gk::k_front /
// gk::k_back.  This file holds the three recurrent kernels, all PERSISTENT over time: a CTA (or thread) owns an utterance / stream,
// walks its frames in order and keeps every recurrent state on chip -- causal conv history in a shared-memory ring, GRU hidden
// state in registers.  Because a GRU has unbounded memory an utterance cannot be split into time segments with a finite warm-up
// (the GTCRN-Micro kernels can): one launch covers frames [0, T) of B utterances, or exactly one frame of N streams when a
// streaming state buffer is given -- same code, same operation order, so streaming is bit-identical to offline.
//
//   kf_gtconv    GTConvBlock: SFE-unfold + 1x1 (24 -> 16), depthwise 3x3 with time dilation d, 1x1 (16 -> 8), TRA (energy ->
//                GRU(8 -> 16) -> Linear -> sigmoid gate), channel shuffle; decoder variant adds the skip.
//   kf_dp_intra  DPGRNN intra path: bidirectional grouped GRU over the 33 bins of a frame, Linear, LayerNorm((33,16)), residual.
//   kf_dp_inter  DPGRNN inter path: grouped GRU over time per (utterance, bin), Linear, LayerNorm((33,16)), residual.
//
// Gate projections run on the FP32 pipe (warp-uniform weights from the constant bank, FFMA2), not on tcgen05: the contractions
// are K = 8 (inputs) and K = 4 / 8 / 16 (hidden) wide with N = 12 / 24 / 48 gate rows, the hidden-side half sits on the serial
// dependency chain of the recurrence, and the 1e-4 waveform tolerance would need a 3xTF32 split.
#pragma once
#include "gtcrn_kernels.cuh"

namespace gf {

using namespace gk;

constexpr int kMaxDil = 5;

// Gate non-linearities on the SFU: ex2.approx + rcp.approx (MUFU.EX2 / MUFU.RCP, ~2^-22 relative error each), 4-5 instructions
// instead of ~20 (expf + IEEE division) and ~30 (tanhf) -- the recurrent kernels evaluate 40 of them per GRU step and were
// issue-bound on exactly that (ncu r2_full_v1: 1,280 warp instructions per inter-GRU step).  The same functions are used by the
// offline and the streaming launches, so the two stay bit-identical; the accuracy cost against the fp32 oracle is ~1e-6.
__device__ __forceinline__ float ex2_approx(float v) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float rcp_approx(float v) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ float sigm(float v) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * v)); }
__device__ __forceinline__ float tanh_(float v) { return fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(2.8853900817779268f * v)), 1.0f); }

// streaming state of the variant: header (same struct as the Micro stream state) + per stream
//   GT ring   6 blocks, 2 d_i rows each (d = 1,2,5,5,2,1 -> 32 rows) x [33][16]   point-conv-1 outputs, slot = frame mod 2d
//   TRA h     6 x 16
//   INTER h   2 x [33][16]                                                      inter-GRU hidden state (group g -> channels 8g..8g+7)
//   WAV_IN 2 x 256, OLA 256                                                     (wav entry points)
constexpr int kHdrFloats = 64;
constexpr int kGtRing = 0;
__host__ __device__ constexpr int gt_ring_row0(int blk) { return blk == 0 ? 0 : blk == 1 ? 2 : blk == 2 ? 6 : blk == 3 ? 16 : blk == 4 ? 26 : 30; }
__host__ __device__ constexpr int gt_dil(int blk) { return blk == 0 ? 1 : blk == 1 ? 2 : blk == 2 ? 5 : blk == 3 ? 5 : blk == 4 ? 2 : 1; }
constexpr int kTraH = 32 * kF2 * 16;                 // 16896
constexpr int kInterH = kTraH + 6 * 16;              // 16992
constexpr int kModelFloats = kInterH + 2 * kF2 * 16; // 18048
constexpr int kWavIn = kModelFloats;
constexpr int kOla = kWavIn + 512;
constexpr int kPerStream = kOla + 256;               // 18816
struct Header { long long frame_step; long long wav_calls; int skip; };
inline size_t state_bytes(int64_t N) { return sizeof(float) * ((size_t)kHdrFloats + (size_t)N * kPerStream); }
__device__ __forceinline__ float* stream_base(void* state, int64_t n) { return reinterpret_cast<float*>(state) + kHdrFloats + n * (int64_t)kPerStream; }

// ================================================================================================
// GTConvBlock (full variant).  grid B (one CTA per utterance / stream), G*33 active threads; a thread owns PS frame slots x 16
// channels at one bin; tiles of NT = G*PS frames.
// ================================================================================================
__host__ __device__ constexpr int gtf_threads(int G) { return (G * kF2 + 31) / 32 * 32; }
constexpr int GTF_GRU_FLOATS = blobf::GTF_SLOPE - blobf::GTF_GRU_WIH;      // W_ih, b_ih, W_hh, b_hh, fc_w, fc_b (padded rows)
__host__ __device__ constexpr int gtf_smem_floats(int PS, int G) {
    const int NT = G * PS;
    return (NT + 2 * kMaxDil) * GT_ROW4 * 4        // ring of point-conv-1 rows
           + NT * 2 * 35 * 4                       // X1: the processed half of the input (8 ch) with zero pad columns, for the unfold
           + NT * 2 * kF2 * 4                      // X2: the untouched half
           + NT * kF2 * 8                          // q^2
           + NT * 8 + NT * 48 + NT * 8             // energies, GRU input projections, gates
           + GTF_GRU_FLOATS;
}

template <bool DEC, int PS, int G>
__global__ void __launch_bounds__(gtf_threads(G), (PS * G <= 1 ? 8 : 3))
kf_gtconv(const float* __restrict__ xin, const float* __restrict__ skip, float* __restrict__ out,
          const __grid_constant__ ParamBlock<blobf::GTF_SIZE> prm, int T, int d, void* state, int blk) {
    constexpr int NT = G * PS, THREADS = gtf_threads(G), ACTIVE = G * kF2;
    const int R = NT + 2 * d;                                         // ring rows in use
    extern __shared__ float4 smem4[];
    float4* Pb = smem4;
    float4* X1 = Pb + (NT + 2 * kMaxDil) * GT_ROW4;                    // [slot][half][35]
    float4* X2 = X1 + NT * 2 * 35;                                     // [slot][half][33]: the untouched half, private to its (slot, bin) thread
    float* E2 = reinterpret_cast<float*>(X2 + NT * 2 * kF2);           // [slot][33][8]
    float* Ev = E2 + NT * kF2 * 8;                                     // [slot][8]
    float* Gi = Ev + NT * 8;                                           // [slot][48]
    float* Gt = Gi + NT * 48;                                          // [slot][8]
    float* Wg = Gt + NT * 8;                                           // GRU / fc weights (per-lane rows: shared memory, padded strides)
    const float* Wih = Wg, *Bih = Wg + (blobf::GTF_GRU_BIH - blobf::GTF_GRU_WIH), *Whh = Wg + (blobf::GTF_GRU_WHH - blobf::GTF_GRU_WIH),
               *Bhh = Wg + (blobf::GTF_GRU_BHH - blobf::GTF_GRU_WIH), *Fcw = Wg + (blobf::GTF_FC_W - blobf::GTF_GRU_WIH), *Fcb = Wg + (blobf::GTF_FC_B - blobf::GTF_GRU_WIH);

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t b = blockIdx.x;
    const Header* hdr = reinterpret_cast<const Header*>(state);
    if (state && hdr->skip) return;
    const long long step0 = state ? hdr->frame_step : 0;              // absolute index of local frame 0 (streaming)
    zero_f4(reinterpret_cast<float*>(smem4), gtf_smem_floats(PS, G) - GTF_GRU_FLOATS, tid, THREADS);
    for (int i = tid; i < GTF_GRU_FLOATS; i += THREADS) Wg[i] = prm.w[blobf::GTF_GRU_WIH + i];
    __syncthreads();                                                  // zeroing done before the history rows are filled in

    const bool act = tid < ACTIVE;
    const int g = tid / kF2, f = tid - g * kF2;
    const int64_t ubase = b * (int64_t)T * kF2 * 16;
    float* sbase = state ? stream_base(state, b) : nullptr;
    float* gring = state ? sbase + kGtRing + gt_ring_row0(blk) * (kF2 * 16) : nullptr;   // [2d][33][16]
    // history rows (local frames -2d .. -1 live in ring rows 0 .. 2d-1): only those a tap of this call's frames will read
    if (state) {
        for (int idx = tid; idx < 2 * d * kF2; idx += THREADS) {
            const int j = idx / kF2, ff = idx - j * kF2;              // local frame j - 2d
            const bool needed = (j - d >= 0 && j - d < T) || (j < T);
            if (!needed) continue;
            long long a = (step0 + j - 2 * d) % (2 * d);
            if (a < 0) a += 2 * d;
            const float4* src = reinterpret_cast<const float4*>(gring + (a * kF2 + ff) * 16);
            const bool before_start = step0 + j - 2 * d < 0;          // zero history before the stream began
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) Pb[j * GT_ROW4 + qd * 35 + ff + 1] = before_start ? make_float4(0.f, 0.f, 0.f, 0.f) : src[qd];
        }
    }
    // TRA GRU state: lanes of warp 0, unit j = lane & 15 (both half-warps carry the same values)
    float hj = 0.f, hall[16];
    if (tid < 32) {
        if (state) hj = sbase[kTraH + blk * 16 + (lane & 15)];
#pragma unroll
        for (int k = 0; k < 16; ++k) hall[k] = __shfl_sync(0xffffffffu, hj, k);
    }
    __syncthreads();

    int rb = 2 * d;                                                   // ring row of the tile's first frame
    for (int t0 = 0; t0 < T; t0 += NT) {
        // ---- phase 0: load x (+ skip); processed half -> X1 (for the 3-bin unfold), untouched half stays in registers
        if (act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                const int t = t0 + g * PS + s;
                float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;
                if (t < T) {
                    const float4* px = reinterpret_cast<const float4*>(xin + ubase + ((int64_t)t * kF2 + f) * 16);
                    v0 = px[0]; v1 = px[1]; v2 = px[2]; v3 = px[3];
                    if (DEC) {
                        const float4* ps = reinterpret_cast<const float4*>(skip + ubase + ((int64_t)t * kF2 + f) * 16);
                        const float4 u0 = ps[0], u1 = ps[1], u2 = ps[2], u3 = ps[3];
                        v0.x += u0.x; v0.y += u0.y; v0.z += u0.z; v0.w += u0.w; v1.x += u1.x; v1.y += u1.y; v1.z += u1.z; v1.w += u1.w;
                        v2.x += u2.x; v2.y += u2.y; v2.z += u2.z; v2.w += u2.w; v3.x += u3.x; v3.y += u3.y; v3.z += u3.z; v3.w += u3.w;
                    }
                }
                X1[((g * PS + s) * 2 + 0) * 35 + f + 1] = v0;
                X1[((g * PS + s) * 2 + 1) * 35 + f + 1] = v1;
                X2[(g * PS + s) * 2 * kF2 + f] = v2;                   // untouched half parked in shared memory until the store phase
                X2[((g * PS + s) * 2 + 1) * kF2 + f] = v3;
            }
        }
        __syncthreads();
        // ---- phase 1: SFE unfold + point conv 1 (24 -> 16) + PReLU -> ring (and the streaming ring in HBM)
        if (act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                const int o = g * PS + s, t = t0 + o;
                float p[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) p[c] = prm.w[blobf::GTF_PC1_B + c];
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 xv = X1[(o * 2 + h) * 35 + f + kk];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int c = 0; c < 16; c += 2)
                                fma_pair(p[c], p[c + 1], f4get(xv, i), prm.w[blobf::GTF_PC1_W + (kk * 8 + 4 * h + i) * 16 + c],
                                         prm.w[blobf::GTF_PC1_W + (kk * 8 + 4 * h + i) * 16 + c + 1]);
                    }
                }
                const float a1 = prm.w[blobf::GTF_SLOPE + 0];
#pragma unroll
                for (int c = 0; c < 16; ++c) p[c] = t < T ? prelu(p[c], a1) : 0.f;
                int row = rb + o; row -= row >= R ? R : 0;
                float4* dst = Pb + row * GT_ROW4 + f + 1;
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) dst[qd * 35] = make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]);
                if (state && t < T && t >= T - 2 * d) {               // newest 2d rows go back to the stream's ring
                    const int a = (int)((step0 + t) % (2 * d));
                    float4* gd = reinterpret_cast<float4*>(gring + ((int64_t)a * kF2 + f) * 16);
#pragma unroll
                    for (int qd = 0; qd < 4; ++qd) gd[qd] = make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]);
                }
            }
        }
        __syncthreads();
        // ---- phase 2: depthwise 3x3 (rows t-2d, t-d, t; bins f-1..f+1) + PReLU, point conv 2 (16 -> 8), q^2
        float q[PS][8];
        if (act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                const int o = g * PS + s, t = t0 + o;
                float dacc[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) dacc[c] = prm.w[blobf::GTF_DW_B + c];
#pragma unroll 1
                for (int kt = 0; kt < 3; ++kt) {                       // rolled: 12 LDS.128 in flight per iteration instead of 36 (spills otherwise)
                    int row = rb + o - (2 - kt) * d;
                    row += row < 0 ? R : 0; row -= row >= R ? R : 0;
                    const float4* src = Pb + row * GT_ROW4 + f;
#pragma unroll
                    for (int kf = 0; kf < 3; ++kf)
#pragma unroll
                        for (int qd = 0; qd < 4; ++qd) {
                            const float4 v = src[qd * 35 + kf];
                            const int wi = blobf::GTF_DW_W + (kt * 3 + kf) * 16 + 4 * qd;
                            fma_pair2(dacc[4 * qd], dacc[4 * qd + 1], v.x, v.y, prm.w[wi], prm.w[wi + 1]);
                            fma_pair2(dacc[4 * qd + 2], dacc[4 * qd + 3], v.z, v.w, prm.w[wi + 2], prm.w[wi + 3]);
                        }
                }
                const float a2 = prm.w[blobf::GTF_SLOPE + 1];
#pragma unroll
                for (int c = 0; c < 16; ++c) dacc[c] = prelu(dacc[c], a2);
#pragma unroll
                for (int c = 0; c < 8; ++c) q[s][c] = prm.w[blobf::GTF_PC2_B + c];
#pragma unroll
                for (int ci = 0; ci < 16; ++ci)
#pragma unroll
                    for (int c = 0; c < 8; c += 2)
                        fma_pair(q[s][c], q[s][c + 1], dacc[ci], prm.w[blobf::GTF_PC2_W + ci * 8 + c], prm.w[blobf::GTF_PC2_W + ci * 8 + c + 1]);
                if (t >= T) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) q[s][c] = 0.f;
                }
                float4* e4 = reinterpret_cast<float4*>(E2 + (o * kF2 + f) * 8);
                e4[0] = make_float4(q[s][0] * q[s][0], q[s][1] * q[s][1], q[s][2] * q[s][2], q[s][3] * q[s][3]);
                e4[1] = make_float4(q[s][4] * q[s][4], q[s][5] * q[s][5], q[s][6] * q[s][6], q[s][7] * q[s][7]);
            }
        }
        __syncthreads();
        // ---- phase 3 (warp 0 only, __syncwarp between its steps): energies, GRU input projections, then the serial GRU walk
        if (tid < 32) {
            for (int i = lane; i < NT * 8; i += 32) Ev[i] = gt_energy(E2 + (i >> 3) * kF2 * 8 + (i & 7));   // fixed summation order
            __syncwarp();
            for (int i = lane; i < NT * 48; i += 32) {
                const int slot = i / 48, row = i - slot * 48;
                float v = Bih[row];
#pragma unroll
                for (int k = 0; k < 8; ++k) v = fmaf(Wih[row * blobf::kWihStride + k], Ev[slot * 8 + k], v);
                Gi[i] = v;
            }
            __syncwarp();
            const int j = lane & 15;
            const int nvalid = min(NT, T - t0);
            const float4* wr = reinterpret_cast<const float4*>(Whh + j * blobf::kWhhStride);
            const float4* wz = reinterpret_cast<const float4*>(Whh + (16 + j) * blobf::kWhhStride);
            const float4* wn = reinterpret_cast<const float4*>(Whh + (32 + j) * blobf::kWhhStride);
            for (int slot = 0; slot < nvalid; ++slot) {
                // three 16-term dot products as 4 partial sums each (dependency depth 4 + 2 instead of 16): this loop is the serial
                // spine of the block -- every other warp of the CTA waits for it
                float pr[4], pz[4], pn[4];
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 a = wr[k4], bq = wz[k4], c = wn[k4];
                    pr[k4] = fmaf(a.w, hall[4 * k4 + 3], fmaf(a.z, hall[4 * k4 + 2], fmaf(a.y, hall[4 * k4 + 1], a.x * hall[4 * k4])));
                    pz[k4] = fmaf(bq.w, hall[4 * k4 + 3], fmaf(bq.z, hall[4 * k4 + 2], fmaf(bq.y, hall[4 * k4 + 1], bq.x * hall[4 * k4])));
                    pn[k4] = fmaf(c.w, hall[4 * k4 + 3], fmaf(c.z, hall[4 * k4 + 2], fmaf(c.y, hall[4 * k4 + 1], c.x * hall[4 * k4])));
                }
                const float gr = __fadd_rn(__fadd_rn(__fadd_rn(pr[0], pr[1]), __fadd_rn(pr[2], pr[3])), Bhh[j]);
                const float gz = __fadd_rn(__fadd_rn(__fadd_rn(pz[0], pz[1]), __fadd_rn(pz[2], pz[3])), Bhh[16 + j]);
                const float gn = __fadd_rn(__fadd_rn(__fadd_rn(pn[0], pn[1]), __fadd_rn(pn[2], pn[3])), Bhh[32 + j]);
                const float r = sigm(__fadd_rn(Gi[slot * 48 + j], gr));
                const float z = sigm(__fadd_rn(Gi[slot * 48 + 16 + j], gz));
                const float n = tanh_(fmaf(r, gn, Gi[slot * 48 + 32 + j]));
                hj = fmaf(z, hj, __fmul_rn(1.0f - z, n));
#pragma unroll
                for (int k = 0; k < 16; ++k) hall[k] = __shfl_sync(0xffffffffu, hj, k);
                if (lane < 16) Gi[slot * 48 + lane] = hj;              // h_t parked in the (consumed) r-gate slots for the Linear below
            }
            __syncwarp();
            // Linear(16 -> 8) + sigmoid for all frames of the tile at once (off the serial spine)
            for (int i = lane; i < nvalid * 8; i += 32) {
                const int slot = i >> 3, c = i & 7;
                float v = Fcb[c];
                const float4* wf = reinterpret_cast<const float4*>(Fcw + c * blobf::kWhhStride);
                const float4* hv = reinterpret_cast<const float4*>(Gi + slot * 48);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 a = wf[k4], hh = hv[k4];
                    v = fmaf(a.x, hh.x, v); v = fmaf(a.y, hh.y, v); v = fmaf(a.z, hh.z, v); v = fmaf(a.w, hh.w, v);
                }
                Gt[i] = sigm(v);
            }
        }
        __syncthreads();
        // ---- phase 4: gate, channel shuffle [q0,x2_0,q1,x2_1,...], store
        if (act) {
#pragma unroll
            for (int s = 0; s < PS; ++s) {
                const int o = g * PS + s, t = t0 + o;
                if (t < T) {
                    const float4 g0 = reinterpret_cast<const float4*>(Gt + o * 8)[0], g1 = reinterpret_cast<const float4*>(Gt + o * 8)[1];
                    const float4 v2 = X2[o * 2 * kF2 + f], v3 = X2[(o * 2 + 1) * kF2 + f];
                    float4* po = reinterpret_cast<float4*>(out + ubase + ((int64_t)t * kF2 + f) * 16);
                    po[0] = make_float4(q[s][0] * g0.x, v2.x, q[s][1] * g0.y, v2.y);
                    po[1] = make_float4(q[s][2] * g0.z, v2.z, q[s][3] * g0.w, v2.w);
                    po[2] = make_float4(q[s][4] * g1.x, v3.x, q[s][5] * g1.y, v3.y);
                    po[3] = make_float4(q[s][6] * g1.z, v3.z, q[s][7] * g1.w, v3.w);
                }
            }
        }
        rb += NT; rb -= rb >= R ? R : 0;
        // the next phase 0 only touches X1 (last read in phase 1, two barriers back); phase 1 then rewrites ring rows whose readers
        // (phase 2 of this tile) are behind the barrier after phase 2
    }
    if (state && tid < 16) sbase[kTraH + blk * 16 + tid] = hj;
}

// ================================================================================================
// DPGRNN intra path.  One CTA = 32 frames x 4 chains: warp w runs chain (group gq = w >> 1, direction = w & 1) for the 32 frames
// (lane = frame), so the GRU weights are warp-uniform (constant bank).  A thread walks the 33 bins of its frame (descending for the
// backward direction), h[4] in registers; outputs land in shared memory [frame][33][16 (+4 pad)] at channel 8 gq + 4 dir.
// Then all 128 threads do Linear(16,16), LayerNorm over the frame's 33 x 16 values (two-pass, fixed order) and the residual add.
// ================================================================================================
constexpr int DPA_FR = 32;
constexpr int DPA_STRIDE = kF2 * 16 + 4;       // 532 floats per frame: lane stride 532 words -> conflict-free LDS.128 / STS.128
constexpr int DPA_SMEM_FLOATS = DPA_FR * DPA_STRIDE + 2 * kF2 * 16 + DPA_FR * 2;
struct DpaParams { float w[blobf::DPA_SIZE]; };

template <int I, int H>   // one GRU step for a chain held by one thread: x[I] inputs, h[H] state; weights [k][3H rows] in a warp-uniform param struct
__device__ __forceinline__ void gru_thread_step(const float* wih, const float* whh, const float* bih, const float* bhh, const float (&x)[I], float (&h)[H]) {
    float gi[3 * H], gh[3 * H];
#pragma unroll
    for (int r = 0; r < 3 * H; ++r) { gi[r] = bih[r]; gh[r] = bhh[r]; }
#pragma unroll
    for (int k = 0; k < I; ++k)
#pragma unroll
        for (int r = 0; r < 3 * H; r += 2) fma_pair(gi[r], gi[r + 1], x[k], wih[k * 3 * H + r], wih[k * 3 * H + r + 1]);
#pragma unroll
    for (int k = 0; k < H; ++k)
#pragma unroll
        for (int r = 0; r < 3 * H; r += 2) fma_pair(gh[r], gh[r + 1], h[k], whh[k * 3 * H + r], whh[k * 3 * H + r + 1]);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const float r = sigm(__fadd_rn(gi[j], gh[j]));
        const float z = sigm(__fadd_rn(gi[H + j], gh[H + j]));
        const float n = tanh_(fmaf(r, gh[2 * H + j], gi[2 * H + j]));
        h[j] = fmaf(z, h[j], __fmul_rn(1.0f - z, n));
    }
}

template <int CH, class PB>
__device__ __forceinline__ void dpa_chain(const PB& prm, const float* __restrict__ xin, float* Y, int lane, int64_t fi0, int64_t nframes) {
    constexpr int gq = CH >> 1, dir = CH & 1;
    const int64_t fi = fi0 + lane;
    const bool valid = fi < nframes;
    const float* xp = xin + (valid ? fi : fi0) * (kF2 * 16) + 8 * gq;
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    const float* wih = prm.w + blobf::DPA_WIH + CH * 96;
    const float* whh = prm.w + blobf::DPA_WHH + CH * 48;
    const float* bih = prm.w + blobf::DPA_BIH + CH * 12;
    const float* bhh = prm.w + blobf::DPA_BHH + CH * 12;
    int f = dir ? kF2 - 1 : 0;
    float4 a = reinterpret_cast<const float4*>(xp + f * 16)[0], bq = reinterpret_cast<const float4*>(xp + f * 16)[1];
#pragma unroll 1
    for (int s = 0; s < kF2; ++s) {
        const int fn = dir ? f - 1 : f + 1;
        float4 an = a, bn = bq;
        if (s + 1 < kF2) { an = reinterpret_cast<const float4*>(xp + fn * 16)[0]; bn = reinterpret_cast<const float4*>(xp + fn * 16)[1]; }   // next bin in flight
        const float x[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
        gru_thread_step<8, 4>(wih, whh, bih, bhh, x, h);
        *reinterpret_cast<float4*>(Y + lane * DPA_STRIDE + f * 16 + 8 * gq + 4 * dir) = make_float4(h[0], h[1], h[2], h[3]);
        a = an; bq = bn; f = fn;
    }
}

__global__ void __launch_bounds__(128, 3)
kf_dp_intra(const float* __restrict__ xin, float* __restrict__ out, const __grid_constant__ ParamBlock<blobf::DPA_SIZE> prm,
            const float* __restrict__ ln /* A_W | A_B, [33][16] each */, int64_t nframes) {
    extern __shared__ float4 smem4[];
    float* Y = reinterpret_cast<float*>(smem4);                 // [32][532]
    float* LnW = Y + DPA_FR * DPA_STRIDE;                        // [528] weight, [528] bias
    float* Stat = LnW + 2 * kF2 * 16;                            // [32][2] mean, rstd
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t fi0 = (int64_t)blockIdx.x * DPA_FR;
    for (int i = tid; i < 2 * kF2 * 16; i += 128) LnW[i] = ln[i];
    switch (w) {                                                 // compile-time chain id keeps the weight indices warp-uniform (LDCU)
        case 0: dpa_chain<0>(prm, xin, Y, lane, fi0, nframes); break;
        case 1: dpa_chain<1>(prm, xin, Y, lane, fi0, nframes); break;
        case 2: dpa_chain<2>(prm, xin, Y, lane, fi0, nframes); break;
        default: dpa_chain<3>(prm, xin, Y, lane, fi0, nframes); break;
    }
    __syncthreads();
    // Linear(16 -> 16) in place: thread = (frame = tid >> 2, r = tid & 3) handles bins r, r+4, ...; per-frame sums via the 4 adjacent lanes
    const int fr = tid >> 2, r = tid & 3;
    float* yf = Y + fr * DPA_STRIDE;
    float sum = 0.f;
    for (int f = r; f < kF2; f += 4) {
        float4 v[4];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) v[qd] = reinterpret_cast<const float4*>(yf + f * 16)[qd];
        float z[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) z[c] = prm.w[blobf::DPA_FC_B + c];
#pragma unroll
        for (int qd = 0; qd < 4; ++qd)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 16; c += 2)
                    fma_pair(z[c], z[c + 1], f4get(v[qd], i), prm.w[blobf::DPA_FC_W + (4 * qd + i) * 16 + c], prm.w[blobf::DPA_FC_W + (4 * qd + i) * 16 + c + 1]);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) reinterpret_cast<float4*>(yf + f * 16)[qd] = make_float4(z[4 * qd], z[4 * qd + 1], z[4 * qd + 2], z[4 * qd + 3]);
#pragma unroll
        for (int c = 0; c < 16; ++c) sum += z[c];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float mean = sum * (1.0f / (kF2 * 16));
    float var = 0.f;
    for (int f = r; f < kF2; f += 4)
#pragma unroll
        for (int c = 0; c < 16; ++c) { const float dlt = yf[f * 16 + c] - mean; var = fmaf(dlt, dlt, var); }
    var += __shfl_xor_sync(0xffffffffu, var, 1);
    var += __shfl_xor_sync(0xffffffffu, var, 2);
    const float rstd = rsqrtf(var * (1.0f / (kF2 * 16)) + 1e-8f);
    if (r == 0) { Stat[2 * fr] = mean; Stat[2 * fr + 1] = rstd; }
    __syncthreads();
    // normalise, affine, residual, store: coalesced over the CTA's contiguous [32][33][16] block
    const int nf = (int)min((int64_t)DPA_FR, nframes - fi0);
    for (int i = tid; i < nf * kF2 * 4; i += 128) {
        const int frm = i / (kF2 * 4), rem = i - frm * (kF2 * 4);          // rem = f*4 + qd
        const float4 z = *reinterpret_cast<const float4*>(Y + frm * DPA_STRIDE + rem * 4);
        const float4 xw = reinterpret_cast<const float4*>(xin + (fi0 + frm) * (kF2 * 16))[rem];
        const float4 lw = reinterpret_cast<const float4*>(LnW)[rem], lb = reinterpret_cast<const float4*>(LnW + kF2 * 16)[rem];
        const float m = Stat[2 * frm], rs = Stat[2 * frm + 1];
        float4 o;
        o.x = xw.x + fmaf((z.x - m) * rs, lw.x, lb.x); o.y = xw.y + fmaf((z.y - m) * rs, lw.y, lb.y);
        o.z = xw.z + fmaf((z.z - m) * rs, lw.z, lb.z); o.w = xw.w + fmaf((z.w - m) * rs, lw.w, lb.w);
        reinterpret_cast<float4*>(out + (fi0 + frm) * (kF2 * 16))[rem] = o;
    }
}

// ================================================================================================
// DPGRNN inter path.  One CTA per utterance / stream, 4 warps: warp 0 / 1 = group 0 / 1 for bins 0..31 (lane = bin), warp 2 / 3 = the
// 33rd bin of group 0 / 1 (one lane) -- a warp only ever runs one group, so the GRU weights stay warp-uniform.  A chain thread keeps
// h[8] in registers across all T frames; every DPE_NT frames the CTA turns the buffered GRU outputs into Linear + LayerNorm +
// residual for those frames.  Streaming: h comes from / returns to the state buffer, T = 1.
// ================================================================================================
constexpr int DPE_NT = 8;
constexpr int DPE_SMEM_FLOATS = DPE_NT * kF2 * 16 + 2 * kF2 * 16 + DPE_NT * 8 + DPE_NT * 2;

template <int GQ, class PB>
__device__ __forceinline__ void dpe_chain(const PB& prm, const float* __restrict__ xp, float* Y, float (&h)[8], int f, int t0, int nv) {
    const float* wih = prm.w + blobf::DPE_WIH + GQ * 192;
    const float* whh = prm.w + blobf::DPE_WHH + GQ * 192;
    const float* bih = prm.w + blobf::DPE_BIH + GQ * 24;
    const float* bhh = prm.w + blobf::DPE_BHH + GQ * 24;
    float4 a = reinterpret_cast<const float4*>(xp + (int64_t)t0 * (kF2 * 16))[0], bq = reinterpret_cast<const float4*>(xp + (int64_t)t0 * (kF2 * 16))[1];
#pragma unroll 1
    for (int s = 0; s < nv; ++s) {
        float4 an = a, bn = bq;
        if (s + 1 < nv) {
            an = reinterpret_cast<const float4*>(xp + (int64_t)(t0 + s + 1) * (kF2 * 16))[0];
            bn = reinterpret_cast<const float4*>(xp + (int64_t)(t0 + s + 1) * (kF2 * 16))[1];
        }
        const float x[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
        gru_thread_step<8, 8>(wih, whh, bih, bhh, x, h);
        float4* y4 = reinterpret_cast<float4*>(Y + (s * kF2 + f) * 16 + 8 * GQ);
        y4[0] = make_float4(h[0], h[1], h[2], h[3]);
        y4[1] = make_float4(h[4], h[5], h[6], h[7]);
        a = an; bq = bn;
    }
}

__global__ void __launch_bounds__(128, 4)
kf_dp_inter(const float* __restrict__ xin, float* __restrict__ out, const __grid_constant__ ParamBlock<blobf::DPE_SIZE> prm,
            const float* __restrict__ ln /* E_W | E_B */, int T, void* state, int which) {
    extern __shared__ float4 smem4[];
    float* Y = reinterpret_cast<float*>(smem4);                 // [DPE_NT][33][16]: GRU outputs, then Linear outputs in place
    float* LnW = Y + DPE_NT * kF2 * 16;
    float* Red = LnW + 2 * kF2 * 16;                             // [DPE_NT][4 warps][2]
    float* Stat = Red + DPE_NT * 8;                              // [DPE_NT][2]
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t b = blockIdx.x;
    const Header* hdr = reinterpret_cast<const Header*>(state);
    if (state && hdr->skip) return;
    for (int i = tid; i < 2 * kF2 * 16; i += 128) LnW[i] = ln[i];
    const int gq = w & 1, f = w < 2 ? lane : kF2 - 1;
    const bool chain = w < 2 || lane == 0;
    const int64_t ubase = b * (int64_t)T * kF2 * 16;
    float* hstate = state ? stream_base(state, b) + kInterH + which * (kF2 * 16) + f * 16 + 8 * gq : nullptr;
    float h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = (state && chain) ? hstate[k] : 0.f;
    const float* xp = xin + ubase + f * 16 + 8 * gq;
    __syncthreads();
    for (int t0 = 0; t0 < T; t0 += DPE_NT) {
        const int nv = min(DPE_NT, T - t0);
        if (chain) {
            if (gq == 0) dpe_chain<0>(prm, xp, Y, h, f, t0, nv);           // compile-time group: warp-uniform weight indices
            else dpe_chain<1>(prm, xp, Y, h, f, t0, nv);
        }
        __syncthreads();
        // Linear(16 -> 16) in place over nv x 33 positions, with per-warp partial sums per frame (fixed order)
        for (int i = tid; i < DPE_NT * kF2; i += 128) {
            const int s = i / kF2;
            if (s >= nv) continue;
            float4 v[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) v[qd] = reinterpret_cast<const float4*>(Y + i * 16)[qd];
            float z[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) z[c] = prm.w[blobf::DPE_FC_B + c];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd)
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int c = 0; c < 16; c += 2)
                        fma_pair(z[c], z[c + 1], f4get(v[qd], k), prm.w[blobf::DPE_FC_W + (4 * qd + k) * 16 + c], prm.w[blobf::DPE_FC_W + (4 * qd + k) * 16 + c + 1]);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) reinterpret_cast<float4*>(Y + i * 16)[qd] = make_float4(z[4 * qd], z[4 * qd + 1], z[4 * qd + 2], z[4 * qd + 3]);
        }
        __syncthreads();
        // LayerNorm statistics per frame: warp s handles frames s, s+4 (two passes over 528 values, lane-strided, shuffle tree)
        for (int s = w; s < nv; s += 4) {
            float sum = 0.f;
            for (int i = lane; i < kF2 * 16; i += 32) sum += Y[s * kF2 * 16 + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float mean = sum * (1.0f / (kF2 * 16));
            float var = 0.f;
            for (int i = lane; i < kF2 * 16; i += 32) { const float dlt = Y[s * kF2 * 16 + i] - mean; var = fmaf(dlt, dlt, var); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
            if (lane == 0) { Stat[2 * s] = mean; Stat[2 * s + 1] = rsqrtf(var * (1.0f / (kF2 * 16)) + 1e-8f); }
        }
        __syncthreads();
        for (int i = tid; i < nv * kF2 * 4; i += 128) {
            const int s = i / (kF2 * 4), rem = i - s * (kF2 * 4);
            const float4 z = reinterpret_cast<const float4*>(Y)[i];
            const int64_t go = ubase + (int64_t)(t0 + s) * (kF2 * 16);
            const float4 xw = reinterpret_cast<const float4*>(xin + go)[rem];
            const float4 lw = reinterpret_cast<const float4*>(LnW)[rem], lb = reinterpret_cast<const float4*>(LnW + kF2 * 16)[rem];
            const float m = Stat[2 * s], rs = Stat[2 * s + 1];
            float4 o;
            o.x = xw.x + fmaf((z.x - m) * rs, lw.x, lb.x); o.y = xw.y + fmaf((z.y - m) * rs, lw.y, lb.y);
            o.z = xw.z + fmaf((z.z - m) * rs, lw.z, lb.z); o.w = xw.w + fmaf((z.w - m) * rs, lw.w, lb.w);
            reinterpret_cast<float4*>(out + go)[rem] = o;
        }
        __syncthreads();
    }
    if (state && chain) {
#pragma unroll
        for (int k = 0; k < 8; ++k) hstate[k] = h[k];
    }
}

}  // namespace gf
