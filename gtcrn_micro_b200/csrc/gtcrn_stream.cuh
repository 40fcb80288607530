// Frame-by-frame (streaming) kernels: one 16 ms hop for N concurrent streams per call.
// reference: StreamGTCRNMicro.forward gtcrn_micro_stream.py:541-574, StreamGTConvBlock :182-262,
//            StreamTCN :265-319, StreamConv2d / StreamConvTranspose2d convolution.py:62-253.
//
// The reference threads shift-register caches through every call (cat([cache, x]) ... cache' = input[:, :, 1:],
// convolution.py:116-119) -- 305 KB of cache traffic per stream and hop.  Here the caches of all N streams live in
// ONE device buffer as RING buffers indexed by the frame counter, so a hop only reads the taps it needs and
// writes the one new frame per cache (~93 KB per stream-hop):
//
//   header (256 B): int64 frame_step, int64 wav_calls, int32 skip
//   per stream (38,880 floats, contiguous):
//     GT_P   [6 blocks][2 slots][33][16]   point-conv-1 outputs of frames t-2, t-1   (blocks 0-2 en2..en4, 3-5 de0..de2)
//     TRA_E  [6 blocks][2 slots][8]        TRALite energies of frames t-2, t-1
//     TCN_Y  [2 gtcn][30 rows][33][16]     conv1 outputs, TCN k owns 2d_k rows starting at row 2^(k+1)-2
//     WAV_IN [2 slots][256]                last two input hops       (wav entry point only)
//     OLA    [256]                         overlap-add tail          (wav entry point only)
//   ring slot of frame t in a cache of depth R: t mod R.
//
// Work mapping: the per-frame kernels of the offline path (k_front, k_back) are reused with T = 1; the temporal
// kernels get tiles ACROSS streams (8 streams x 33 bins per CTA for GTConv, 512 (stream,bin) columns per CTA for
// the two GTCNs, which are purely per-column recurrences and need no shared-memory exchange at all).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "blob_layout.h"
#include "gtcrn_kernels.cuh"

namespace gk {
// defined in gtcrn_tc.cuh: one decoder GTConvBlock of a streaming hop on the tensor cores
inline cudaError_t launch_dec_tc_stream(const float* x, const float* skip, float* out, const ParamBlock<gt_chunk<true>()>& prm, void* state, int blk,
                                        int64_t N, int sm_count, cudaStream_t st);
}

namespace gs {

using gk::kF1;
using gk::kF2;

constexpr int kHdrFloats = 64;
constexpr int kGtP = 0;
constexpr int kTraE = kGtP + 6 * 2 * kF2 * 16;            // 6336
constexpr int kTcnY = kTraE + 6 * 2 * 8;                  // 6432
constexpr int kModelFloats = kTcnY + 2 * 30 * kF2 * 16;   // 38112
constexpr int kWavIn = kModelFloats;
constexpr int kOla = kWavIn + 2 * 256;
constexpr int kPerStream = kOla + 256;                    // 38880

struct Header {
    long long frame_step;
    long long wav_calls;
    int skip;
};

struct Ws {          // per-hop activation workspace (owned by the handle), all [N][F][16]
    float* en0;
    float* act[9];
    float* spec_in;
    float* spec_out;
};

inline size_t state_bytes(int64_t N) { return sizeof(float) * ((size_t)kHdrFloats + (size_t)N * kPerStream); }

// where the wav-streaming tails of a stream live (the streaming STFT / iSTFT kernels serve both model variants)
struct WavLayout { int per_stream, wav_in, ola; };
constexpr WavLayout kWavLayoutMicro{kPerStream, kWavIn, kOla};
__device__ __forceinline__ float* wav_base(void* state, int64_t n, const WavLayout& wl) {
    return reinterpret_cast<float*>(state) + kHdrFloats + n * (int64_t)wl.per_stream;
}

__device__ __forceinline__ float* stream_base(void* state, int64_t n) {
    return reinterpret_cast<float*>(state) + kHdrFloats + n * (int64_t)kPerStream;
}

// ------------------------------------------------------------------------------------------------
// GTConvBlock, one frame for SG_NS streams per CTA.  Same device functions and operation order as gk::k_gtconv
// (bit-identical results); the two history rows of every stream come from the GT_P ring instead of the previous tile.
// grid ceil(N/8), 288 threads (264 active = 8 streams x 33 bins), one stream per thread.
// ------------------------------------------------------------------------------------------------
constexpr int SG_NS = 8;
constexpr int SG_THREADS = 288;
constexpr int SG_ACTIVE = SG_NS * kF2;
constexpr int SG_ROW4 = gk::GT_ROW4;
constexpr int SG_SMEM_FLOATS = SG_NS * 3 * SG_ROW4 * 4 + SG_NS * kF2 * 8 + SG_NS * 3 * 8 + SG_NS * 8;

template <bool DEC>
__global__ void __launch_bounds__(SG_THREADS, DEC ? 2 : 3)
ks_gtconv(const float* __restrict__ xin, const float* __restrict__ skip, float* __restrict__ out,
          const __grid_constant__ gk::ParamBlock<gk::gt_chunk<DEC>()> prm, void* state, int blk, int64_t N) {
    using namespace gk;
    const Header* hdr = reinterpret_cast<const Header*>(state);
    if (hdr->skip) return;
    const int slot_old = (int)(hdr->frame_step & 1);          // holds frame t-2; overwritten with frame t
    const int slot_mid = slot_old ^ 1;                        // holds frame t-1

    extern __shared__ float4 smem4[];
    float4* Pb = smem4;                                               // [stream][3 rows: t-2, t-1, t][q][col]
    float* E2 = reinterpret_cast<float*>(Pb + SG_NS * 3 * SG_ROW4);   // [stream][33][8]
    float* Ev = E2 + SG_NS * kF2 * 8;                                 // [stream][3][8]
    float* Gt = Ev + SG_NS * 3 * 8;                                   // [stream][8]

    const int tid = threadIdx.x;
    const int64_t n0 = (int64_t)blockIdx.x * SG_NS;
    // only the pad columns (0 and 34 of every row) and the rows of absent streams must read as zero: everything else is written
    // before it is read (zeroing all 53 KB per CTA was 5.5 % of the kernel's samples, ncu r2_stream16k_v19)
    for (int i = tid; i < SG_NS * 3 * 4 * 2; i += SG_THREADS) Pb[(i >> 1) * 35 + (i & 1) * 34] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + SG_NS > N) {
        zero_f4(reinterpret_cast<float*>(Pb), SG_SMEM_FLOATS, tid, SG_THREADS);
        __syncthreads();
    }

    const int ls = tid / kF2, f = tid - ls * kF2;
    const int64_t n = n0 + ls;
    const bool act = tid < SG_ACTIVE && n < N;
    float* ring = nullptr;
    // ---- phase 1: history rows from the ring, point conv 1 of the new frame (also written to the ring)
    if (act) {
        ring = stream_base(state, n) + kGtP + blk * (2 * kF2 * 16) + f * 16;
        const float4* h2 = reinterpret_cast<const float4*>(ring + slot_old * (kF2 * 16));
        const float4* h1 = reinterpret_cast<const float4*>(ring + slot_mid * (kF2 * 16));
        float4* row = Pb + (ls * 3) * SG_ROW4 + f + 1;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) { row[qd * 35] = h2[qd]; row[SG_ROW4 + qd * 35] = h1[qd]; }
        const float4* px = reinterpret_cast<const float4*>(xin + (n * kF2 + f) * 16);
        float4 v0 = px[0], v1 = px[1];
        if (DEC) {
            const float4* ps = reinterpret_cast<const float4*>(skip + (n * kF2 + f) * 16);
            const float4 u0 = ps[0], u1 = ps[1];
            v0.x += u0.x; v0.y += u0.y; v0.z += u0.z; v0.w += u0.w;
            v1.x += u1.x; v1.y += u1.y; v1.z += u1.z; v1.w += u1.w;
        }
        const float x1[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float p[16];
        gt_pc1<DEC>(prm, x1, p);
        float4* newest = reinterpret_cast<float4*>(ring + slot_old * (kF2 * 16));
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float4 v = make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]);
            newest[qd] = v;
            row[2 * SG_ROW4 + qd * 35] = v;
        }
    }
    if (tid < SG_NS * 16) {                                           // TRA energy history -> Ev rows 0,1
        const int l2 = tid >> 4, r = (tid >> 3) & 1, c = tid & 7;
        const int64_t n2 = n0 + l2;
        float v = 0.f;
        if (n2 < N) v = stream_base(state, n2)[kTraE + blk * 16 + (r == 0 ? slot_old : slot_mid) * 8 + c];
        Ev[(l2 * 3 + r) * 8 + c] = v;
    }
    __syncthreads();
    // ---- phase 2: 3x3 conv + PReLU, point conv 2, energies
    float q[8];
    if (act) {
        float acc[1][16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[0][c] = prm.w[GtOff<DEC>::MID_B + c];
        const float4* r0 = Pb + (ls * 3) * SG_ROW4 + f;
        // kf rolled, ci-quads unrolled (same shape as the offline kernel): keeps the weight loads on the uniform datapath
        // (UR-indexed LDCU) instead of per-thread-indexed LDC through the indexed constant cache
#pragma unroll 1
        for (int kf = 0; kf < 3; ++kf) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float4 in[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) in[j] = r0[j * SG_ROW4 + qd * 35 + kf];
                gt_mid_step<DEC, 1>(prm, kf, qd, in, acc);
            }
        }
        gt_pc2<DEC>(prm, acc[0], q);
        float4* e4 = reinterpret_cast<float4*>(E2 + (ls * kF2 + f) * 8);
        e4[0] = make_float4(q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]);
        e4[1] = make_float4(q[4] * q[4], q[5] * q[5], q[6] * q[6], q[7] * q[7]);
    }
    __syncthreads();
    // ---- phase 3: TRALite energy of the new frame (same summation order as the offline kernel) -> ring
    if (tid < SG_NS * 8) {
        const int l2 = tid >> 3, c = tid & 7;
        const float e = gt_energy(E2 + l2 * kF2 * 8 + c);
        Ev[(l2 * 3 + 2) * 8 + c] = e;
        const int64_t n2 = n0 + l2;
        if (n2 < N) stream_base(state, n2)[kTraE + blk * 16 + slot_old * 8 + c] = e;
    }
    __syncthreads();
    if (tid < SG_NS * 8) {
        const int l2 = tid >> 3, c = tid & 7;
        Gt[l2 * 8 + c] = gt_gate1<DEC>(prm, Ev + (l2 * 3) * 8, Ev + (l2 * 3 + 1) * 8, Ev + (l2 * 3 + 2) * 8, c);
    }
    __syncthreads();
    // ---- phase 4: gate, shuffle, store
    if (act) {
        const float4* px = reinterpret_cast<const float4*>(xin + (n * kF2 + f) * 16);
        float4 v2 = px[2], v3 = px[3];
        if (DEC) {
            const float4* ps = reinterpret_cast<const float4*>(skip + (n * kF2 + f) * 16);
            const float4 u2 = ps[2], u3 = ps[3];
            v2.x += u2.x; v2.y += u2.y; v2.z += u2.z; v2.w += u2.w;
            v3.x += u3.x; v3.y += u3.y; v3.z += u3.z; v3.w += u3.w;
        }
        const float* gate = Gt + ls * 8;
        float4* po = reinterpret_cast<float4*>(out + (n * kF2 + f) * 16);
        po[0] = make_float4(q[0] * gate[0], v2.x, q[1] * gate[1], v2.y);
        po[1] = make_float4(q[2] * gate[2], v2.z, q[3] * gate[3], v2.w);
        po[2] = make_float4(q[4] * gate[4], v3.x, q[5] * gate[5], v3.y);
        po[3] = make_float4(q[6] * gate[6], v3.z, q[7] * gate[7], v3.w);
    }
}

// ------------------------------------------------------------------------------------------------
// Both GTCNs (8 TCN blocks) for one frame: a pure per-(stream, bin) recurrence.  A thread owns two columns (every weight feeds
// two FFMA2), keeps x in registers through all eight TCNs and touches exactly 2 taps + 1 new row per TCN in the ring.
// The tap addresses depend only on the frame counter, so the taps of TCN k+1 are requested with cp.async (LDGSTS: global ->
// shared, no staging registers) before TCN k is computed and read back from a thread-private, lane-interleaved shared-memory
// slot: the ring reads -- the bulk of a hop's HBM traffic -- overlap the previous TCN's 1,024 FFMA2 per thread.  [The round-1
// kernel loaded the taps right before use: 61 % of its stall samples sat on the first tap, 1.96 TB/s at 16,384 streams,
// profiles/ncu_r2_stream16k_v19_lines_ks_gtcn.txt; holding the next taps in registers instead made ptxas fall back to
// per-thread-indexed LDC for the weights.]  Weights: kernel parameter / constant bank.  grid ceil(N*33 / 256), 128 threads.
// ------------------------------------------------------------------------------------------------
constexpr int ST_THREADS = 128;
constexpr int ST_SMEM_BYTES = 2 * 16 * ST_THREADS * 16;          // 2 stages x 16 float4 per thread

__global__ void __launch_bounds__(ST_THREADS, 3)
ks_gtcn(const float* __restrict__ xin, float* __restrict__ out, const __grid_constant__ gk::ParamBlock<8 * blob::TCN_SIZE> prm,
        void* state, int64_t N) {
    using namespace gk;
    extern __shared__ float4 smem4[];                             // [stage][item = (col, tap, quad)][thread]
    const Header* hdr = reinterpret_cast<const Header*>(state);
    if (hdr->skip) return;
    const long long step = hdr->frame_step;
    const int tid = threadIdx.x;
    const int64_t npos = N * kF2;
    int64_t pos[2] = {(int64_t)blockIdx.x * (2 * ST_THREADS) + tid, (int64_t)blockIdx.x * (2 * ST_THREADS) + tid + ST_THREADS};
    bool ok[2] = {pos[0] < npos, pos[1] < npos};
    if (!ok[0]) return;                                      // pos[1] > pos[0]: nothing to do for this thread
    float* ybase[2];
    float x[2][16];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const int64_t p = ok[s] ? pos[s] : pos[0];           // an inactive second column mirrors the first (never stored)
        const int64_t n = p / kF2;
        const int f = (int)(p - n * kF2);
        ybase[s] = stream_base(state, n) + kTcnY + f * 16;
    }
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(smem4 + tid);
    auto prefetch = [&](int gk_) {                           // taps of TCN gk_ -> stage gk_ & 1
        const int g = gk_ >> 2, k = gk_ & 3, d = 1 << k, row0 = g * 30 + (2 << k) - 2;
        const int slotA = (int)(step & (2 * d - 1)), slotB = (int)((step + d) & (2 * d - 1));
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const float* rA = ybase[s] + (row0 + slotA) * (kF2 * 16);
            const float* rB = ybase[s] + (row0 + slotB) * (kF2 * 16);
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                const unsigned da = sm0 + (unsigned)((((gk_ & 1) * 16 + s * 8 + qd) * ST_THREADS) * 16);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(rA + 4 * qd) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + 4u * ST_THREADS * 16u), "l"(rB + 4 * qd) : "memory");
            }
        }
        cp_async_commit();
    };
    prefetch(0);
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const float4* px = reinterpret_cast<const float4*>(xin + (ok[s] ? pos[s] : pos[0]) * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float4 u = px[v];
            x[s][4 * v] = u.x; x[s][4 * v + 1] = u.y; x[s][4 * v + 2] = u.z; x[s][4 * v + 3] = u.w;
        }
    }
#pragma unroll 1
    for (int gk_ = 0; gk_ < 8; ++gk_) {
        const int g = gk_ >> 2, k = gk_ & 3, d = 1 << k, row0 = g * 30 + (2 << k) - 2, wo = gk_ * blob::TCN_SIZE;
        const int slotA = (int)(step & (2 * d - 1));          // frame t-2d, overwritten with frame t
        if (gk_ < 7) prefetch(gk_ + 1); else cp_async_commit();
        float y1[2][16];
        tcn_conv1<2>(prm, wo, x, y1);
        cp_async_wait_group<1>();                             // this TCN's taps have landed (the next one's may still be in flight)
        const float4* st4 = smem4 + ((gk_ & 1) * 16) * ST_THREADS + tid;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            float4 ta[4], tb[4];
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) { ta[qd] = st4[(s * 8 + qd) * ST_THREADS]; tb[qd] = st4[(s * 8 + 4 + qd) * ST_THREADS]; }
            if (s == 0 || ok[1]) {
                float4* rA = reinterpret_cast<float4*>(ybase[s] + (row0 + slotA) * (kF2 * 16));
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) rA[qd] = make_float4(y1[s][4 * qd], y1[s][4 * qd + 1], y1[s][4 * qd + 2], y1[s][4 * qd + 3]);
            }
            tcn_dw(prm, wo, ta, tb, y1[s]);
        }
        tcn_conv3<2>(prm, wo, y1, x);
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (!ok[s]) continue;
        float4* po = reinterpret_cast<float4*>(out + pos[s] * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) po[v] = make_float4(x[s][4 * v], x[s][4 * v + 1], x[s][4 * v + 2], x[s][4 * v + 3]);
    }
}

// ------------------------------------------------------------------------------------------------
// streaming STFT / iSTFT for the wav entry point.  Call k (0-based) receives input hop k.
//   k = 0 : nothing can be framed yet (frame 0 needs sample 256): hop stored, model skipped, zeros out
//   k >= 1: frame k-1 = [hop k-2 | hop k-1]  (frame 0 = [reflect(hop 0, x[256]) | hop 0]) -> model -> iSTFT;
//           output hop k-2 = OLA tail of frame k-2 + first half of frame k-1 (zeros for k = 1: trimmed pre-roll)
// so the enhanced stream lags the input by two hops (512 samples); every emitted hop equals the offline output.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ks_stft_step(void* state, const float* __restrict__ hop_in, const float* __restrict__ blobp,
                                                    float* __restrict__ spec, int64_t N, WavLayout wl) {
    using namespace gk;
    __shared__ FftTables tb;
    __shared__ __align__(16) float2 buf[4][2][256];
    Header* hdr = reinterpret_cast<Header*>(state);
    const long long k = hdr->wav_calls;
    const int tid = threadIdx.x, fr = tid >> 6, j = tid & 63;
    const int64_t n = (int64_t)blockIdx.x * 4 + fr;
    const bool valid = n < N;
    load_fft_tables(tb, blobp, tid, 256);
    __syncthreads();
    float2* A = buf[fr][0];
    float2* Bf = buf[fr][1];
    float* ring = valid ? wav_base(state, n, wl) + wl.wav_in : nullptr;
    const float* nh = valid ? hop_in + n * 256 : nullptr;
    if (valid && k >= 1) {
        const float* h_prev = ring + ((k - 1) & 1) * 256;      // hop k-1
        const float* h_prev2 = ring + (k & 1) * 256;           // hop k-2 (garbage for k = 1, unused then)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int m = j + 64 * r;
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int i = 2 * m + e;                       // position in the 512-sample frame
                if (i >= 256) v[e] = h_prev[i - 256];
                else if (k >= 2) v[e] = h_prev2[i];
                else v[e] = (256 - i == 256) ? nh[0] : h_prev[256 - i];     // reflect: x[256 - i], x[256] is the new hop's first sample
            }
            A[m] = make_float2(tb.win[2 * m] * v[0], tb.win[2 * m + 1] * v[1]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) A[j + 64 * r] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (valid) {                                              // store hop k over hop k-2 (all reads of it are done)
#pragma unroll
        for (int r = 0; r < 4; ++r) ring[(k & 1) * 256 + j + 64 * r] = nh[j + 64 * r];
    }
    fft256(A, Bf, tb.tw256, j);
    if (valid && k >= 1) {
        float2* o = reinterpret_cast<float2*>(spec + n * kSpecFrame);
        for (int kk = j; kk <= 256; kk += 64) {
            const float2 zk = A[kk & 255];
            float2 zc = A[(256 - kk) & 255];
            zc.y = -zc.y;
            const float2 w = tb.tw512[kk];
            const float dx = zk.x - zc.x, dy = zk.y - zc.y;
            o[kk] = make_float2(0.5f * (zk.x + zc.x) + 0.5f * (w.x * dy + w.y * dx),
                                0.5f * (zk.y + zc.y) - 0.5f * (w.x * dx - w.y * dy));
        }
    }
    if (blockIdx.x == 0 && tid == 0) hdr->skip = (k == 0) ? 1 : 0;
}

__global__ void __launch_bounds__(256) ks_istft_step(void* state, const float* __restrict__ spec, const float* __restrict__ blobp,
                                                     float* __restrict__ hop_out, int64_t N, WavLayout wl) {
    using namespace gk;
    __shared__ FftTables tb;
    __shared__ __align__(16) float2 buf[4][2][256];
    __shared__ float inv_env[256];
    const Header* hdr = reinterpret_cast<const Header*>(state);
    const long long k = hdr->wav_calls;
    const int tid = threadIdx.x, fr = tid >> 6, j = tid & 63;
    const int64_t n = (int64_t)blockIdx.x * 4 + fr;
    const bool valid = n < N;
    load_fft_tables(tb, blobp, tid, 256);
    inv_env[tid] = blobp[blob::FFT_OFF + blob::FFT_INV_ENV + tid];
    __syncthreads();
    float2* A = buf[fr][0];
    float2* Bf = buf[fr][1];
    if (valid && k >= 1) {
        const float* xt = spec + n * kSpecFrame;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kk = j + 64 * r;
            float2 xk = make_float2(xt[2 * kk], xt[2 * kk + 1]);
            float2 xc = make_float2(xt[2 * (256 - kk)], -xt[2 * (256 - kk) + 1]);
            if (kk == 0) { xk.y = 0.f; xc.y = 0.f; }
            const float2 w = tb.tw512[kk];
            const float dx = xk.x - xc.x, dy = xk.y - xc.y;
            const float a = w.x * dx + w.y * dy, bq = w.x * dy - w.y * dx;
            A[kk] = make_float2(0.5f * (xk.x + xc.x) - 0.5f * bq, -(0.5f * (xk.y + xc.y) + 0.5f * a));
        }
    } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) A[j + 64 * r] = make_float2(0.f, 0.f);
    }
    __syncthreads();
    fft256(A, Bf, tb.tw256, j);
    if (!valid) return;
    float* o = hop_out + n * 256;
    if (k == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r) o[j + 64 * r] = 0.f;
        return;
    }
    float* tail = wav_base(state, n, wl) + wl.ola;
    // time-domain samples 2m, 2m+1 of the frame = (Re z[m], -Im z[m]) / 256, windowed.  m = j + 64 r: the thread that
    // consumes tail[2m] (r = 0,1) is the same thread that rewrites it from sample m + 128 (r = 2,3) -> no barrier needed.
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int m = j + 64 * r;
        const float2 z = A[m];
        const float v0 = tb.win[2 * m] * (z.x * (1.0f / 256.0f)), v1 = tb.win[2 * m + 1] * (-z.y * (1.0f / 256.0f));
        if (m < 128) {                                        // first half: completes output hop k-2
            const float2 t2 = *reinterpret_cast<const float2*>(tail + 2 * m);
            const float2 res = (k >= 2) ? make_float2((t2.x + v0) * inv_env[2 * m], (t2.y + v1) * inv_env[2 * m + 1]) : make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(o + 2 * m) = res;
        } else {                                              // second half: new overlap-add tail
            *reinterpret_cast<float2*>(tail + 2 * (m - 128)) = make_float2(v0, v1);
        }
    }
}

// End of stream (gtcrnb_stream_flush_wav): after K full hops (and `tail_len` < 256 trailing samples) the offline path still has
// two frames to go -- frame K-1 = [hop K-2 | hop K-1] and frame K = [hop K-1 | v], v = the "virtual hop K" that torch.stft's
// reflect padding creates past the end of the signal (infer.py:60-67; L = 256 K + tail_len, x[i >= L] = x[2 (L-1) - i]).
// This kernel builds v [N][256] from the trailing samples and the two hops in the WAV_IN ring; two ordinary step_wav calls
// with v as the new hop then emit the last two offline hops.
__global__ void __launch_bounds__(256) ks_reflect_tail(const void* state, const float* __restrict__ tail, int tail_len,
                                                       float* __restrict__ v, int64_t N, WavLayout wl) {
    const Header* hdr = reinterpret_cast<const Header*>(state);
    const long long K = hdr->wav_calls;
    const int64_t n = blockIdx.x;
    const int j = threadIdx.x;
    if (n >= N) return;
    const float* ring = reinterpret_cast<const float*>(state) + kHdrFloats + n * (int64_t)wl.per_stream + wl.wav_in;
    const float* h1 = ring + ((K - 1) & 1) * 256;              // hop K-1
    const float* h2 = ring + (K & 1) * 256;                    // hop K-2
    float val = 0.f;
    if (K >= 2) {
        if (j < tail_len) val = tail[n * tail_len + j];
        else {
            const int r = 2 * tail_len - 2 - j;                // position relative to sample 256 K
            val = r >= 0 ? tail[n * tail_len + r] : (r >= -256 ? h1[256 + r] : h2[512 + r]);
        }
    }
    v[n * 256 + j] = val;
}

__global__ void ks_advance(void* state, int wav_mode) {
    Header* hdr = reinterpret_cast<Header*>(state);
    if (!hdr->skip) hdr->frame_step += 1;
    if (wav_mode) hdr->wav_calls += 1;
    hdr->skip = 0;
}

// ------------------------------------------------------------------------------------------------
// import / export of the model caches in the reference's tensor layout (gtcrn_micro_stream.py:618-623)
//   conv_cache [2][N][16][6][33], tra_cache [2][3][N][8][2], tcn_cache[g][k] [N][16][2d][33]
// export: j-th oldest entry of a depth-R cache = ring slot (step + j) mod R.  import: step := 0, slot j := entry j.
// ------------------------------------------------------------------------------------------------
struct TcnPtrs { float* p[8]; };

__global__ void __launch_bounds__(256) ks_convert(int import, void* state, int64_t N, float* conv_cache, float* tra_cache, TcnPtrs tcn) {
    Header* hdr = reinterpret_cast<Header*>(state);
    const long long step = import ? 0 : hdr->frame_step;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * kModelFloats) return;
    const int64_t n = idx / kModelFloats;
    const int e = (int)(idx - n * kModelFloats);
    float* mine = stream_base(state, n) + e;
    float* theirs;
    if (e < kTraE) {                                          // GT_P [blk][slot][f][c]
        const int blk = e / (2 * kF2 * 16), r = e - blk * (2 * kF2 * 16);
        const int slot = r / (kF2 * 16), f = (r / 16) % kF2, c = r & 15;
        const int j = (int)((slot - step) & 1);               // age index (0 = oldest) of this ring slot
        const int half = blk / 3, i = blk % 3;
        const int fr = (half == 0 ? 2 * i : 4 - 2 * i) + j;   // decoder blocks use the slices in reverse (de0 -> [4:6])
        theirs = conv_cache + ((((int64_t)half * N + n) * 16 + c) * 6 + fr) * kF2 + f;
    } else if (e < kTcnY) {                                   // TRA_E [blk][slot][c]
        const int r = e - kTraE, blk = r / 16, slot = (r >> 3) & 1, c = r & 7;
        const int j = (int)((slot - step) & 1);
        const int half = blk / 3, i = blk % 3;
        theirs = tra_cache + ((((int64_t)half * 3 + i) * N + n) * 8 + c) * 2 + j;
    } else {                                                  // TCN_Y [g][30 rows][f][c]
        const int r = e - kTcnY, g = r / (30 * kF2 * 16), rr = r - g * (30 * kF2 * 16);
        const int row = rr / (kF2 * 16), f = (rr / 16) % kF2, c = rr & 15;
        const int k = row < 2 ? 0 : row < 6 ? 1 : row < 14 ? 2 : 3;
        const int d = 1 << k, slot = row - ((2 << k) - 2);
        const int j = (int)((slot - step) & (2 * d - 1));
        theirs = tcn.p[g * 4 + k] + (((int64_t)n * 16 + c) * (2 * d) + j) * kF2 + f;
    }
    if (import) *mine = *theirs; else *theirs = *mine;
    if (import && idx == 0) { hdr->frame_step = 0; hdr->skip = 0; }
}

// ------------------------------------------------------------------------------------------------
// The three encoder GTConvBlocks of a hop in ONE launch (round 2): a thread keeps its position's 16 channels in registers from
// block to block (the channel shuffle is a register permutation), every block's output is still written once for the decoder
// skip, and the shared-memory tile is reused.  Same device functions and operation order as ks_gtconv<false> x 3 (bit-identical);
// saves two launches and the re-read of two activation tensors per hop (ncu r2_stream16k_v23: the three launches were 182 us of
// an 826 us hop at 16,384 streams, 23 % issue utilisation, stalled on barriers and on the loads at the head of each kernel).
// ------------------------------------------------------------------------------------------------
template <class PB>
__device__ __forceinline__ void sg_enc_block(const PB& prm, void* state, int blk, int64_t N, int64_t n0, int tid, bool act, int ls, int f, int64_t n,
                                             int slot_old, int slot_mid, float4* Pb, float* E2, float* Ev, float* Gt, float (&x)[16],
                                             float* __restrict__ out) {
    using namespace gk;
    // ---- phase 1: history rows from the ring, point conv 1 of the new frame (also written to the ring)
    if (act) {
        float* ring = stream_base(state, n) + kGtP + blk * (2 * kF2 * 16) + f * 16;
        const float4* h2 = reinterpret_cast<const float4*>(ring + slot_old * (kF2 * 16));
        const float4* h1 = reinterpret_cast<const float4*>(ring + slot_mid * (kF2 * 16));
        float4* row = Pb + (ls * 3) * SG_ROW4 + f + 1;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) { row[qd * 35] = h2[qd]; row[SG_ROW4 + qd * 35] = h1[qd]; }
        const float x1[8] = {x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]};
        float p[16];
        gt_pc1<false>(prm, x1, p);
        float4* newest = reinterpret_cast<float4*>(ring + slot_old * (kF2 * 16));
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
            const float4 v = make_float4(p[4 * qd], p[4 * qd + 1], p[4 * qd + 2], p[4 * qd + 3]);
            newest[qd] = v;
            row[2 * SG_ROW4 + qd * 35] = v;
        }
    }
    if (tid < SG_NS * 16) {                                           // TRA energy history -> Ev rows 0,1
        const int l2 = tid >> 4, r = (tid >> 3) & 1, c = tid & 7;
        const int64_t n2 = n0 + l2;
        float v = 0.f;
        if (n2 < N) v = stream_base(state, n2)[kTraE + blk * 16 + (r == 0 ? slot_old : slot_mid) * 8 + c];
        Ev[(l2 * 3 + r) * 8 + c] = v;
    }
    __syncthreads();
    // ---- phase 2: 3x3 depthwise conv + PReLU, point conv 2, energies
    float q[8];
    if (act) {
        float acc[1][16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[0][c] = prm.w[GtOff<false>::MID_B + c];
        const float4* r0 = Pb + (ls * 3) * SG_ROW4 + f;
#pragma unroll 1
        for (int kf = 0; kf < 3; ++kf) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
                float4 in[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) in[j] = r0[j * SG_ROW4 + qd * 35 + kf];
                gt_mid_step<false, 1>(prm, kf, qd, in, acc);
            }
        }
        gt_pc2<false>(prm, acc[0], q);
        float4* e4 = reinterpret_cast<float4*>(E2 + (ls * kF2 + f) * 8);
        e4[0] = make_float4(q[0] * q[0], q[1] * q[1], q[2] * q[2], q[3] * q[3]);
        e4[1] = make_float4(q[4] * q[4], q[5] * q[5], q[6] * q[6], q[7] * q[7]);
    }
    __syncthreads();
    // ---- phase 3: TRALite energy of the new frame -> ring, gate
    if (tid < SG_NS * 8) {
        const int l2 = tid >> 3, c = tid & 7;
        const float e = gt_energy(E2 + l2 * kF2 * 8 + c);
        Ev[(l2 * 3 + 2) * 8 + c] = e;
        const int64_t n2 = n0 + l2;
        if (n2 < N) stream_base(state, n2)[kTraE + blk * 16 + slot_old * 8 + c] = e;
    }
    __syncthreads();
    if (tid < SG_NS * 8) {
        const int l2 = tid >> 3, c = tid & 7;
        Gt[l2 * 8 + c] = gt_gate1<false>(prm, Ev + (l2 * 3) * 8, Ev + (l2 * 3 + 1) * 8, Ev + (l2 * 3 + 2) * 8, c);
    }
    __syncthreads();
    // ---- phase 4: gate, channel shuffle (in registers: the next block's input), store
    if (act) {
        const float* gate = Gt + ls * 8;
        float y[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) { y[2 * i] = q[i] * gate[i]; y[2 * i + 1] = x[8 + i]; }
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = y[c];
        float4* po = reinterpret_cast<float4*>(out + (n * kF2 + f) * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) po[v] = make_float4(x[4 * v], x[4 * v + 1], x[4 * v + 2], x[4 * v + 3]);
    }
}

__global__ void __launch_bounds__(SG_THREADS, 3)
ks_gtconv_enc3(const float* __restrict__ xin, float* __restrict__ out1, float* __restrict__ out2, float* __restrict__ out3,
               const __grid_constant__ gk::ParamBlock<blob::ENCGT_SIZE> prm0, const __grid_constant__ gk::ParamBlock<blob::ENCGT_SIZE> prm1,
               const __grid_constant__ gk::ParamBlock<blob::ENCGT_SIZE> prm2, void* state, int64_t N) {
    using namespace gk;
    const Header* hdr = reinterpret_cast<const Header*>(state);
    if (hdr->skip) return;
    const int slot_old = (int)(hdr->frame_step & 1), slot_mid = slot_old ^ 1;
    extern __shared__ float4 smem4[];
    float4* Pb = smem4;
    float* E2 = reinterpret_cast<float*>(Pb + SG_NS * 3 * SG_ROW4);
    float* Ev = E2 + SG_NS * kF2 * 8;
    float* Gt = Ev + SG_NS * 3 * 8;
    const int tid = threadIdx.x;
    const int64_t n0 = (int64_t)blockIdx.x * SG_NS;
    for (int i = tid; i < SG_NS * 3 * 4 * 2; i += SG_THREADS) Pb[(i >> 1) * 35 + (i & 1) * 34] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + SG_NS > N) {
        zero_f4(reinterpret_cast<float*>(Pb), SG_SMEM_FLOATS, tid, SG_THREADS);
        __syncthreads();
    }
    const int ls = tid / kF2, f = tid - ls * kF2;
    const int64_t n = n0 + ls;
    const bool act = tid < SG_ACTIVE && n < N;
    float x[16];
    if (act) {
        const float4* px = reinterpret_cast<const float4*>(xin + (n * kF2 + f) * 16);
#pragma unroll
        for (int v = 0; v < 4; ++v) { const float4 u = px[v]; x[4 * v] = u.x; x[4 * v + 1] = u.y; x[4 * v + 2] = u.z; x[4 * v + 3] = u.w; }
        // the ring rows of the two later blocks start their trip from HBM now
        const float* r1 = stream_base(state, n) + kGtP + 1 * (2 * kF2 * 16) + f * 16;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + kF2 * 16));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + 2 * kF2 * 16));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + 3 * kF2 * 16));
    } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = 0.f;
    }
    sg_enc_block(prm0, state, 0, N, n0, tid, act, ls, f, n, slot_old, slot_mid, Pb, E2, Ev, Gt, x, out1);
    sg_enc_block(prm1, state, 1, N, n0, tid, act, ls, f, n, slot_old, slot_mid, Pb, E2, Ev, Gt, x, out2);
    sg_enc_block(prm2, state, 2, N, n0, tid, act, ls, f, n, slot_old, slot_mid, Pb, E2, Ev, Gt, x, out3);
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
inline int stream_init_attributes() {
    cudaError_t e;
    const void* fns[] = {(const void*)ks_gtconv<false>, (const void*)ks_gtconv<true>, (const void*)ks_stft_step, (const void*)ks_istft_step};
    for (const void* f : fns) {
        e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaFuncSetAttribute(ks_gtcn, cudaFuncAttributeMaxDynamicSharedMemorySize, ST_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ks_gtcn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ks_gtconv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_FLOATS * (int)sizeof(float));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ks_gtconv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_FLOATS * (int)sizeof(float));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ks_gtconv_enc3, cudaFuncAttributeMaxDynamicSharedMemorySize, SG_SMEM_FLOATS * (int)sizeof(float));
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(ks_gtconv_enc3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    return (int)e;
}

// host copies of the per-kernel weight chunks, passed by value as kernel parameters (constant bank)
struct HostParams {
    gk::ParamBlock<blob::ENCGT_SIZE> enc[3];
    gk::ParamBlock<blob::DECGT_SIZE> dec[3];
    gk::ParamBlock<4 * blob::TCN_SIZE> gtcn[2];
    gk::ParamBlock<8 * blob::TCN_SIZE> tcn8;
    gk::ParamBlock<blob::FRONT_SIZE> front;
    gk::ParamBlock<blob::BACK_SIZE> back;
    bool tc_stream = false;                       // decoder blocks of a hop on the tensor cores (k_gtconv_dec_tc<4, true>)
    bool fuse_enc = true;                         // the three encoder GTConvBlocks of a hop in one launch (GTCRNB_STREAM_FUSE_ENC=0: three)
    bool tc_ends = false;                         // en1 on the tensor cores (k_front<.., true>), weights in tcw
    bool tc_back = false;                         // de3 on the tensor cores (k_back<.., true>)
    const uint32_t* tcw = nullptr;
    int sm_count = 148;
    void load(const float* blob_host) {
        memcpy(front.w, blob_host + blob::FRONT_OFF, sizeof(front.w));
        memcpy(back.w, blob_host + blob::BACK_OFF, sizeof(back.w));
        for (int i = 0; i < 3; ++i) {
            memcpy(enc[i].w, blob_host + blob::ENCGT_OFF + i * blob::ENCGT_SIZE, sizeof(enc[i].w));
            memcpy(dec[i].w, blob_host + blob::DECGT_OFF + i * blob::DECGT_SIZE, sizeof(dec[i].w));
        }
        for (int i = 0; i < 2; ++i) memcpy(gtcn[i].w, blob_host + blob::TCN_OFF + i * 4 * blob::TCN_SIZE, sizeof(gtcn[i].w));
        memcpy(tcn8.w, blob_host + blob::TCN_OFF, sizeof(tcn8.w));
    }
};

// model step on spectra already in `spec` ([N][257][2] contiguous); returns the number of launches
inline int model_step(const float* blobp, const HostParams& hp, void* state, const Ws& ws, const float* spec, float* out, int64_t N, cudaStream_t st) {
    using namespace gk;
    int launches = 0;
    const unsigned gfr = (unsigned)((N + FR_NF - 1) / FR_NF), gbk = (unsigned)((N + BK_NF - 1) / BK_NF);
    // T = 1: strides {b: 514, f: 2, t: 0, c: 1}.  k_front / k_back do not know about `skip`; on the skipped call
    // (first wav hop) they run on a zero spectrum and their outputs are simply never consumed.
    if (hp.tc_ends)
        k_front<false, 0, true><<<gfr, FR_THREADS, fr_smem_floats(false, true) * sizeof(float), st>>>(spec, kSpecFrame, 2, 0, 1, nullptr, 0, 0, nullptr, 1, N, blobp, hp.front,
                                                                                                      ws.en0, ws.act[0], hp.tcw);
    else
        k_front<false, 0><<<gfr, FR_THREADS, fr_smem_floats(false) * sizeof(float), st>>>(spec, kSpecFrame, 2, 0, 1, nullptr, 0, 0, nullptr, 1, N, blobp, hp.front, ws.en0, ws.act[0]);
    ++launches;
    const unsigned gsg = (unsigned)((N + SG_NS - 1) / SG_NS);
    if (hp.fuse_enc) {
        ks_gtconv_enc3<<<gsg, SG_THREADS, SG_SMEM_FLOATS * sizeof(float), st>>>(ws.act[0], ws.act[1], ws.act[2], ws.act[3], hp.enc[0], hp.enc[1], hp.enc[2], state, N);
        ++launches;
    } else
    for (int i = 0; i < 3; ++i) {
        ks_gtconv<false><<<gsg, SG_THREADS, SG_SMEM_FLOATS * sizeof(float), st>>>(ws.act[i], nullptr, ws.act[i + 1], hp.enc[i], state, i, N);
        ++launches;
    }
    ks_gtcn<<<(unsigned)((N * kF2 + 2 * ST_THREADS - 1) / (2 * ST_THREADS)), ST_THREADS, ST_SMEM_BYTES, st>>>(ws.act[3], ws.act[5], hp.tcn8, state, N);
    ++launches;
    for (int i = 0; i < 3; ++i) {
        if (hp.tc_stream) gk::launch_dec_tc_stream(ws.act[5 + i], ws.act[3 - i], ws.act[6 + i], hp.dec[i], state, 3 + i, N, hp.sm_count, st);
        else ks_gtconv<true><<<gsg, SG_THREADS, SG_SMEM_FLOATS * sizeof(float), st>>>(ws.act[5 + i], ws.act[3 - i], ws.act[6 + i], hp.dec[i], state, 3 + i, N);
        ++launches;
    }
    if (hp.tc_back)
        k_back<false, 0, true><<<gbk, BK_THREADS, bk_smem_floats(false, true) * sizeof(float), st>>>(ws.act[8], ws.act[0], ws.en0, spec, kSpecFrame, 2, 0, 1,
                                                                                                     out, kSpecFrame, 2, 0, 1, nullptr, 0, 1, N, blobp, hp.back, hp.tcw);
    else
    k_back<false, 0><<<gbk, BK_THREADS, bk_smem_floats(false) * sizeof(float), st>>>(ws.act[8], ws.act[0], ws.en0, spec, kSpecFrame, 2, 0, 1,
                                                                                  out, kSpecFrame, 2, 0, 1, nullptr, 0, 1, N, blobp, hp.back);
    ++launches;
    return launches;
}

inline int step_spec(const float* blobp, const HostParams& hp, void* state, const Ws& ws, const float* spec, float* out, int64_t N, cudaStream_t st) {
    int launches = model_step(blobp, hp, state, ws, spec, out, N, st);
    ks_advance<<<1, 1, 0, st>>>(state, 0);
    return launches + 1;
}

inline int step_wav(const float* blobp, const HostParams& hp, void* state, const Ws& ws, const float* hop_in, float* hop_out, int64_t N, cudaStream_t st) {
    const unsigned g4 = (unsigned)((N + 3) / 4);
    ks_stft_step<<<g4, 256, 0, st>>>(state, hop_in, blobp, ws.spec_in, N, kWavLayoutMicro);
    int launches = 1 + model_step(blobp, hp, state, ws, ws.spec_in, ws.spec_out, N, st);
    ks_istft_step<<<g4, 256, 0, st>>>(state, ws.spec_out, blobp, hop_out, N, kWavLayoutMicro);
    ks_advance<<<1, 1, 0, st>>>(state, 1);
    return launches + 2;
}

// the last two offline hops; out [2][N][256] (out[1] doubles as the buffer of the virtual hop).  Returns the number of launches.
inline int flush_wav(const float* blobp, const HostParams& hp, void* state, const Ws& ws, const float* tail, int tail_len, float* out, int64_t N, cudaStream_t st) {
    float* v = out + N * 256;
    ks_reflect_tail<<<(unsigned)N, 256, 0, st>>>(state, tail, tail_len, v, N, kWavLayoutMicro);
    int launches = 1 + step_wav(blobp, hp, state, ws, v, out, N, st);
    launches += step_wav(blobp, hp, state, ws, v, v, N, st);
    return launches;
}

inline int convert_caches(bool import, void* state, int64_t N, float* conv_cache, float* tra_cache, float* const* tcn, cudaStream_t st) {
    TcnPtrs tp;
    for (int i = 0; i < 8; ++i) tp.p[i] = tcn[i];
    const int64_t total = N * kModelFloats;
    ks_convert<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(import ? 1 : 0, state, N, conv_cache, tra_cache, tp);
    return 1;
}

}  // namespace gs
