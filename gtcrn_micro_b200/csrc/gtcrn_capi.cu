// C-ABI of libgtcrn_b200.so (declared in include/gtcrn_b200.h): handle, activation workspace,
// kernel launch plumbing.  No torch types, no CPU fallback: every entry point either launches the
// sm_100a kernels of gtcrn_kernels.cuh / gtcrn_stream.cuh or returns an error code.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "../../include/gtcrn_b200.h"
#include "gtcrn_kernels.cuh"
#include "gtcrn_stream.cuh"
#include "gtcrn_tc.cuh"
#include "gtcrn_full_host.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(GTCRNB_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kAbiVersion = 1;
constexpr int64_t kMaxChunkFrames = 1024 * 1024;  // activation workspace capacity (frames): 1024 x 10 s utterances fit one chunk (17.5 GB)
constexpr int kNumAct33 = 9;                      // en1..en4, gtcn1, gtcn2, de0..de2
const char* const kActNames[kNumAct33] = {"en1", "en2", "en3", "en4", "gtcn1", "gtcn2", "de0", "de1", "de2"};

enum KernelClass { KC_STFT = 0, KC_FRONT, KC_ENC_GT, KC_GTCN, KC_DEC_GT, KC_BACK, KC_ISTFT, KC_STREAM,
                   KC_FULL_ENC_GT, KC_DP_INTRA, KC_DP_INTER, KC_FULL_DEC_GT, kNumKernelClasses };
struct ProfEvent { int cls; int64_t frames; cudaEvent_t e0, e1; };

}  // namespace

struct gtcrnb_handle {
    int device = 0;
    int sm_count = 148;
    float* blob = nullptr;
    int variant = 0;                              // 0 GTCRN-Micro, 1 original-GTCRN-style (BASELINE configs[3])
    gs::HostParams* hp = nullptr;                 // host copies of the weight chunks passed as kernel parameters
    gf::HostParams* hpf = nullptr;                // same for variant 1
    bool fuse_fft = true;                         // fuse STFT / iSTFT into the front / back kernels in gtcrnb_enhance
    bool tcn_smem_w = false;                      // GTCN weights from shared memory (LDS.128) instead of the constant bank; with FFMA2 the
                                                  // constant-bank feed wins (B200: 4.73 vs 5.67 ms per 1024 x 10 s), before FFMA2 smem won (6.4 vs 7.3)
    int tcn_p = 2;                                // frame slots per lane of the GTCN kernel (2..4)
    // tensor-core (tcgen05, 3xTF32) kernels: bit 0 GTCN pointwise convs, bit 1 decoder GTConv 3x3 (GTCRNB_TC, default both);
    // 0 selects the FFMA kernels, whose results are bit-identical to the streaming path
    int tc_mask = 15;                             // bit 2: en1 (k_front), bit 3: de3 (k_back) as tcgen05 MMAs (round 2)
    uint32_t* tcw = nullptr;                      // their fp16 hi / lo weight matrices (k_prep_ends_tc)
    int tc_gtcn_nwg = 1;                          // warpgroups per CTA of k_gtcn_tc (GTCRNB_TC_GTCN_NWG: 1, 2 or 4)
    int occ_tcn_tc = 1;
    int tc_dec_nwg = 4;                           // warpgroups per CTA of k_gtconv_dec_tc (GTCRNB_TC_DEC_NWG: 1, 2, 4 or 5; 96 TMEM columns each)
    bool tc_stack = false;                        // [W_hi | W_lo] stacked along N: 4 MMAs per conv and 32 accumulator columns instead of 6 and 16
    int gt_enc = 0, gt_dec = 0;                   // index into kGtVariants for the encoder / decoder GTConv kernels
    int seg_waves = 1;                            // CTAs per resident slot of the walking kernels (GTCRNB_SEG_WAVES), see pick_flat
    // activation workspace for one chunk of utterances
    int64_t cap_frames = 0;
    float* en0 = nullptr;
    float* act[kNumAct33] = {};
    float* spec_in = nullptr;
    float* spec_out = nullptr;
    // host pipeline staging
    static constexpr int kStage = 3;                 // utterance groups in flight: staging buffers per direction and activation workspaces
    float* stage_in[kStage] = {};
    float* stage_out[kStage] = {};
    size_t stage_in_floats = 0, stage_out_floats = 0;
    int16_t* pcm_in[kStage] = {};                    // PCM16 wire buffers of gtcrnb_enhance_host_pcm16
    int16_t* pcm_out[kStage] = {};
    size_t pcm_in_elems = 0, pcm_out_elems = 0;
    // host pipeline streams: copies, per-copy-chunk front / back kernels, per-group walking kernels (see host_pipeline)
    cudaStream_t s_h2d = nullptr, s_front = nullptr, s_body = nullptr, s_body2 = nullptr, s_back = nullptr, s_d2h = nullptr;
    // activation workspaces: set 0 serves the caller-stream entry points, the host pipeline rotates through all kNumWs (one per
    // utterance group in flight).  The fields above are the currently selected set; sets[ws_cur] is stale while it is selected.
    struct WsSet { int64_t cap_frames = 0; float* en0 = nullptr; float* act[kNumAct33] = {}; float* spec_in = nullptr; float* spec_out = nullptr; };
    static constexpr int kNumWs = kStage;
    WsSet sets[kNumWs];
    int ws_cur = 0;
    cudaEvent_t ev_fronts[kStage] = {}, ev_body[kStage] = {}, ev_backs[kStage] = {}, ev_d2h[kStage] = {};
    std::vector<cudaEvent_t> ev_pool;             // per-copy-chunk events of the host pipeline (grown on demand, reused across calls)
    // Ordering between the caller's stream and the host pipeline's internal streams: both use the handle's activation workspace.
    // Every entry point that launches on a caller stream records ev_user there; the host pipeline makes its compute streams wait
    // on it before touching the workspace (the pipeline itself is synchronous, so nothing is needed in the other direction).
    cudaEvent_t ev_user = nullptr;
    bool user_pending = false;
    bool in_pipeline = false;
    // occupancy (resident CTAs per SM) of the walking kernels, for segment selection
    int occ_gt_enc = 1, occ_gt_dec = 1, occ_tcn = 1;
    int64_t launches = 0;
    int64_t last_B = 0, last_T = 0;
    // optional per-kernel-class timing with CUDA events on the launching stream (gtcrnb_profile_*)
    bool profiling = false;
    // CUDA graphs of the streaming hop (10-12 kernels whose by-value weight blocks, up to 20 KB each, make the hop launch-bound
    // at small stream counts): instantiated the second consecutive time the same (state, in, out, N) is stepped, replayed after
    struct StepGraph { int kind; const void* state; const void* in; const void* out; int64_t n; cudaGraphExec_t exec; int launches; uint64_t used; };
    std::vector<StepGraph> step_graphs;
    std::vector<StepGraph> seen_steps;               // recent hops that ran uncaptured (keys only)
    cudaStream_t s_cap = nullptr;
    uint64_t step_tick = 0;
    bool stream_graphs = true;                       // GTCRNB_STREAM_GRAPH=0 disables
    std::vector<ProfEvent> prof_events;
    double prof_ms[kNumKernelClasses] = {};
    int64_t prof_launches[kNumKernelClasses] = {};
    int64_t prof_frames[kNumKernelClasses] = {};
};

namespace {

void drop_step_graphs(gtcrnb_handle* h) {
    for (auto& g : h->step_graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    h->step_graphs.clear();
    h->seen_steps.clear();
}

int ensure_workspace(gtcrnb_handle* h, int64_t frames) {
    if (frames <= h->cap_frames) return GTCRNB_OK;
    drop_step_graphs(h);                              // cached hop graphs point into the old workspace
    auto freep = [](float*& p) { if (p) cudaFree(p); p = nullptr; };
    freep(h->en0); freep(h->spec_in); freep(h->spec_out);
    for (auto& p : h->act) freep(p);
    h->cap_frames = 0;
    auto alloc = [&](float*& p, size_t floats_per_frame) -> bool {
        return cudaMalloc(&p, sizeof(float) * floats_per_frame * (size_t)frames) == cudaSuccess;
    };
    bool ok = alloc(h->en0, gk::kF1 * 16) && alloc(h->spec_in, gk::kSpecFrame) && alloc(h->spec_out, gk::kSpecFrame);
    for (auto& p : h->act) ok = ok && alloc(p, gk::kF2 * 16);
    if (!ok) {
        cudaGetLastError();
        return fail(GTCRNB_E_NOMEM, "workspace allocation for %lld frames failed", (long long)frames);
    }
    h->cap_frames = frames;
    return GTCRNB_OK;
}

// GTConv walking-kernel variants (frame slots per thread PS, slot groups per CTA G)
struct GtVariant { int ps, g, smem_w, threads, smem_enc, smem_dec; const void* enc; const void* dec; };
template <int PS, int G, bool SW> GtVariant make_gt_variant() {
    return {PS, G, SW ? 1 : 0, gk::gt_threads(G),
            gk::gt_smem_floats(PS, G, SW ? gk::gt_chunk<false>() : 0) * (int)sizeof(float),
            gk::gt_smem_floats(PS, G, SW ? gk::gt_chunk<true>() : 0) * (int)sizeof(float),
            (const void*)gk::k_gtconv<false, PS, G, SW>, (const void*)gk::k_gtconv<true, PS, G, SW>};
}
// index 0-2: (2,8) (3,8) (4,7) constant bank; 3-5: the same with shared-memory weights; 6..: further constant-bank shapes
const GtVariant kGtVariants[] = {make_gt_variant<2, 8, false>(), make_gt_variant<3, 8, false>(), make_gt_variant<4, 7, false>(),
                                 make_gt_variant<2, 8, true>(), make_gt_variant<3, 8, true>(), make_gt_variant<4, 7, true>(),
                                 make_gt_variant<1, 8, false>(), make_gt_variant<2, 4, false>()};
// Encoder: (2,4) -- 8-frame tiles, 160 threads, 5 CTAs/SM -- is the default: 2.84 ms per 1024 x 10 s against 3.03 for (2,8); the
// kernel is latency / barrier bound, so more and smaller CTAs win [(2,3) 2.86, (2,6) 2.90, (2,5) 2.94, (4,4) 3.03, (3,4) 3.18].
// Decoder: FMA bound, wants the largest tile that keeps two CTAs per SM: (4,7) 7.51 ms [(4,5) 8.30, (4,4) 9.47, (3,8) 8.6].
constexpr int kNumGtVariants = (int)(sizeof(kGtVariants) / sizeof(kGtVariants[0]));

template <bool DEC>
cudaError_t launch_gtconv(const GtVariant& v, dim3 grid, cudaStream_t st, const float* x, const float* skip, float* out,
                          const gk::ParamBlock<gk::gt_chunk<DEC>()>& prm, int T, int seg, int B) {
    void* args[] = {(void*)&x, (void*)&skip, (void*)&out, (void*)&prm, (void*)&T, (void*)&seg, (void*)&B};
    return cudaLaunchKernel(DEC ? v.dec : v.enc, grid, dim3((unsigned)v.threads), args, (size_t)(DEC ? v.smem_dec : v.smem_enc), st);
}

const void* tc_dec_fn(int nwg) {
    using namespace gk;
    return nwg == 1 ? (const void*)k_gtconv_dec_tc<1, false> : nwg == 2 ? (const void*)k_gtconv_dec_tc<2, false>
         : nwg == 4 ? (const void*)k_gtconv_dec_tc<4, false> : (const void*)k_gtconv_dec_tc<5, false>;
}
const void* tc_gtcn_fn(int nwg, bool stack) {
    using namespace gk;
    if (stack) return nwg == 1 ? (const void*)k_gtcn_tc<1, true> : nwg == 2 ? (const void*)k_gtcn_tc<2, true> : (const void*)k_gtcn_tc<4, true>;
    return nwg == 1 ? (const void*)k_gtcn_tc<1, false> : nwg == 2 ? (const void*)k_gtcn_tc<2, false> : (const void*)k_gtcn_tc<4, false>;
}

// RAII: brackets a kernel launch with two events when profiling is on
struct ProfScope {
    gtcrnb_handle* h; cudaStream_t st; ProfEvent ev; bool on;
    ProfScope(gtcrnb_handle* h_, int cls, int64_t frames, cudaStream_t st_) : h(h_), st(st_), on(h_->profiling) {
        if (!on) return;
        ev.cls = cls; ev.frames = frames;
        cudaEventCreate(&ev.e0); cudaEventCreate(&ev.e1);
        cudaEventRecord(ev.e0, st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(ev.e1, st);
        h->prof_events.push_back(ev);
    }
};

// called after launching on a caller-owned stream (see gtcrnb_handle::ev_user)
void note_user_work(gtcrnb_handle* h, cudaStream_t st) {
    if (h->in_pipeline) return;
    if (!h->ev_user && cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); h->ev_user = nullptr; return; }
    if (cudaEventRecord(h->ev_user, st) == cudaSuccess) h->user_pending = true; else cudaGetLastError();
}

void select_workspace(gtcrnb_handle* h, int idx) {
    if (h->ws_cur == idx) return;
    gtcrnb_handle::WsSet& cur = h->sets[h->ws_cur];
    cur.cap_frames = h->cap_frames; cur.en0 = h->en0; cur.spec_in = h->spec_in; cur.spec_out = h->spec_out;
    for (int i = 0; i < kNumAct33; ++i) cur.act[i] = h->act[i];
    const gtcrnb_handle::WsSet& nxt = h->sets[idx];
    h->cap_frames = nxt.cap_frames; h->en0 = nxt.en0; h->spec_in = nxt.spec_in; h->spec_out = nxt.spec_out;
    for (int i = 0; i < kNumAct33; ++i) h->act[i] = nxt.act[i];
    h->ws_cur = idx;
}

// Split T frames into `nseg` equal segments for a walking kernel: minimise waves * (segment work + warm-up + per-CTA overhead).
// Segments need not be tile aligned (tiles start at the segment's own warm-up frame and slot groups past the end are skipped);
// `gran` is the granularity at which trailing work is skipped (frames).
int pick_seg_len(int64_t ctas_per_seg_row, int T, int gran, int warm, int slots) {
    double best = 1e300;
    int best_len = T;
    const int max_seg = std::max(1, std::min(64, T / (4 * warm + 1) + 1));
    for (int nseg = 1; nseg <= max_seg; ++nseg) {
        const int seg_len = (T + nseg - 1) / nseg;
        const int64_t ctas = ctas_per_seg_row * nseg;
        const double waves = (double)((ctas + slots - 1) / slots);
        const double work = (double)((seg_len + (nseg > 1 ? warm : 0) + gran - 1) / gran * gran) + 8.0;   // + fixed per-CTA prologue
        const double cost = waves * work;
        if (cost < best - 1e-9) { best = cost; best_len = seg_len; }
    }
    return best_len;
}

// Balanced segmentation of a walking kernel: `units` x T frames flattened, one contiguous range of `len` frames per CTA, `waves`
// CTAs per resident slot (1 = every CTA does exactly 1 / slots of the work).  A range is never shorter than 4 warm-ups.
struct FlatSeg { int len; unsigned ctas; };
FlatSeg pick_flat(int64_t units, int T, int warm, int64_t slots, int waves) {
    const int64_t total = units * (int64_t)T, want = std::max<int64_t>(1, slots * std::max(1, waves));
    int64_t len = std::max<int64_t>((total + want - 1) / want, std::min<int64_t>(total, 4 * warm));
    len = std::min<int64_t>(len, (int64_t)1 << 30);
    return {(int)len, (unsigned)((total + len - 1) / len)};
}

// Phases of the model, so that the host pipeline can run the per-frame ends (front: STFT .. en1, back: de3 .. iSTFT) on small
// copy chunks -- utterances [u0, u0 + nu) of the workspace -- and the walking kernels in between on whole utterance groups.
enum { PH_FRONT = 1, PH_BODY = 2, PH_BACK = 4, PH_ALL = 7 };

template <int VAR, class HP>
void launch_front(gtcrnb_handle* h, const HP& hp, const float* spec, const int64_t is[4], int64_t T, int64_t u0, int64_t nu, cudaStream_t st,
                  const float* wav, int64_t wav_stride, int L) {
    using namespace gk;
    const int64_t nframes = nu * T, f0 = u0 * T;
    const dim3 gfr((unsigned)((nframes + FR_NF - 1) / FR_NF));
    ProfScope ps(h, KC_FRONT, nframes, st);
    float* en0 = h->en0 + f0 * kF1 * 16;
    float* en1 = h->act[0] + f0 * kF2 * 16;
    if constexpr (VAR == 0) {
        if (h->tc_mask & 4) {
            if (wav)
                k_front<true, 0, true><<<gfr, FR_THREADS, fr_smem_floats(true, true) * sizeof(float), st>>>(nullptr, 0, 0, 0, 0, wav + u0 * wav_stride, wav_stride, L,
                                                                                                            h->spec_in + f0 * kSpecFrame, (int)T, nframes, h->blob, hp.front, en0, en1, h->tcw);
            else
                k_front<false, 0, true><<<gfr, FR_THREADS, fr_smem_floats(false, true) * sizeof(float), st>>>(spec + u0 * is[0], is[0], is[1], is[2], is[3], nullptr, 0, 0, nullptr,
                                                                                                              (int)T, nframes, h->blob, hp.front, en0, en1, h->tcw);
            h->launches++;
            return;
        }
    }
    if (wav)
        k_front<true, VAR><<<gfr, FR_THREADS, fr_smem_floats(true) * sizeof(float), st>>>(nullptr, 0, 0, 0, 0, wav + u0 * wav_stride, wav_stride, L,
                                                                                           h->spec_in + f0 * kSpecFrame, (int)T, nframes, h->blob, hp.front, en0, en1);
    else
        k_front<false, VAR><<<gfr, FR_THREADS, fr_smem_floats(false) * sizeof(float), st>>>(spec + u0 * is[0], is[0], is[1], is[2], is[3], nullptr, 0, 0, nullptr,
                                                                                             (int)T, nframes, h->blob, hp.front, en0, en1);
    h->launches++;
}

template <int VAR, class HP>
void launch_back(gtcrnb_handle* h, const HP& hp, const float* spec, const int64_t is[4], int64_t T, int64_t u0, int64_t nu, float* out, const int64_t os[4],
                 cudaStream_t st, bool fused, float* wav_out, int64_t wav_out_stride) {
    using namespace gk;
    const int64_t nframes = nu * T, f0 = u0 * T;
    const dim3 gbk((unsigned)((nframes + BK_NF - 1) / BK_NF));
    ProfScope ps(h, KC_BACK, nframes, st);
    const float* de2 = h->act[8] + f0 * kF2 * 16;
    const float* en1 = h->act[0] + f0 * kF2 * 16;
    const float* en0 = h->en0 + f0 * kF1 * 16;
    const int64_t ss[4] = {T * kSpecFrame, 2, kSpecFrame, 1};
    if constexpr (VAR == 0) {
        if (h->tc_mask & 8) {
            if (fused)
                k_back<true, 0, true><<<gbk, BK_THREADS, bk_smem_floats(true, true) * sizeof(float), st>>>(de2, en1, en0, h->spec_in + f0 * kSpecFrame, ss[0], ss[1], ss[2], ss[3],
                                                                                                           nullptr, 0, 0, 0, 0, wav_out + u0 * wav_out_stride, wav_out_stride,
                                                                                                           (int)T, nframes, h->blob, hp.back, h->tcw);
            else
                k_back<false, 0, true><<<gbk, BK_THREADS, bk_smem_floats(false, true) * sizeof(float), st>>>(de2, en1, en0, spec + u0 * is[0], is[0], is[1], is[2], is[3],
                                                                                                             out + u0 * os[0], os[0], os[1], os[2], os[3], nullptr, 0,
                                                                                                             (int)T, nframes, h->blob, hp.back, h->tcw);
            h->launches++;
            return;
        }
    }
    if (fused) {
        k_back<true, VAR><<<gbk, BK_THREADS, bk_smem_floats(true) * sizeof(float), st>>>(de2, en1, en0, h->spec_in + f0 * kSpecFrame, ss[0], ss[1], ss[2], ss[3],
                                                                                          nullptr, 0, 0, 0, 0, wav_out + u0 * wav_out_stride, wav_out_stride,
                                                                                          (int)T, nframes, h->blob, hp.back);
    } else {
        k_back<false, VAR><<<gbk, BK_THREADS, bk_smem_floats(false) * sizeof(float), st>>>(de2, en1, en0, spec + u0 * is[0], is[0], is[1], is[2], is[3],
                                                                                            out + u0 * os[0], os[0], os[1], os[2], os[3], nullptr, 0,
                                                                                            (int)T, nframes, h->blob, hp.back);
    }
    h->launches++;
}

// the walking kernels of GTCRN-Micro on B whole utterances: 3 encoder GTConv blocks, 2 GTCNs, 3 decoder GTConv blocks
void launch_body_micro(gtcrnb_handle* h, int64_t B, int64_t T, cudaStream_t st) {
    using namespace gk;
    const int Ti = (int)T;
    const int64_t nframes = B * T;
    {
        const GtVariant& v = kGtVariants[h->gt_enc];
        const FlatSeg fs = pick_flat(B, Ti, GT_WARMUP, (int64_t)h->sm_count * h->occ_gt_enc, h->seg_waves);
        const dim3 g(fs.ctas);
        for (int i = 0; i < 3; ++i) {
            ProfScope ps(h, KC_ENC_GT, nframes, st);
            launch_gtconv<false>(v, g, st, h->act[i], nullptr, h->act[i + 1], h->hp->enc[i], Ti, fs.len, (int)B);
            h->launches++;
        }
    }
    {
        const int64_t ncols = B * kF2;
        const int64_t ctas = (ncols + WC_WARPS - 1) / WC_WARPS;
        const int seg = pick_seg_len(ctas, Ti, 32, TC_WARMUP, h->sm_count * h->occ_tcn);
        const dim3 g((unsigned)ctas, (unsigned)((Ti + seg - 1) / seg));
        if (h->tc_mask & 1) {
            const int nwg = h->tc_gtcn_nwg;
            const int64_t ctas_tc = (ncols + 4 * nwg - 1) / (4 * nwg);
            const int seg_tc = pick_seg_len(ctas_tc, Ti, 32, TC_WARMUP, h->sm_count * h->occ_tcn_tc);
            const dim3 gt((unsigned)ctas_tc, (unsigned)((Ti + seg_tc - 1) / seg_tc));
            for (int i = 0; i < 2; ++i) {
                ProfScope ps(h, KC_GTCN, nframes, st);
                const size_t sm = tct_smem_floats(nwg) * sizeof(float);
                const float* xi = h->act[3 + i]; float* xo = h->act[4 + i];
                int seg_i = seg_tc; int64_t nc = ncols; int Tl = Ti;
                void* args[] = {(void*)&xi, (void*)&xo, (void*)&h->hp->gtcn[i], (void*)&Tl, (void*)&seg_i, (void*)&nc};
                cudaLaunchKernel(tc_gtcn_fn(nwg, h->tc_stack), gt, dim3(128 * nwg), args, sm, st);
                h->launches++;
            }
        } else
        for (int i = 0; i < 2; ++i) {
            ProfScope ps(h, KC_GTCN, nframes, st);
            if (h->tcn_smem_w) k_gtcn<2, true><<<g, WC_WARPS * 32, wc_smem_floats(2, true) * sizeof(float), st>>>(h->act[3 + i], h->act[4 + i], h->hp->gtcn[i], Ti, seg, ncols);
            else k_gtcn<2, false><<<g, WC_WARPS * 32, wc_smem_floats(2) * sizeof(float), st>>>(h->act[3 + i], h->act[4 + i], h->hp->gtcn[i], Ti, seg, ncols);
            h->launches++;
        }
    }
    if (h->tc_mask & 2) {
        const int nwg = h->tc_dec_nwg;
        const int rows = (int)((B + GTT_G * nwg - 1) / (GTT_G * nwg));
        const int seg = pick_seg_len(rows, Ti, 1, GT_WARMUP, h->sm_count * gtt_min_blocks(nwg));
        const dim3 g((unsigned)((Ti + seg - 1) / seg), (unsigned)rows);
        for (int i = 0; i < 3; ++i) {
            ProfScope ps(h, KC_DEC_GT, nframes, st);
            const float* xi = h->act[5 + i]; const float* sk = h->act[3 - i]; float* xo = h->act[6 + i];
            int seg_i = seg, Bi = (int)B, Tl = Ti;
            void* nostate = nullptr; int blk0 = 0; long long n0 = 0;
#ifdef GTCRNB_DEC_TRACE
            static long long* trc_dev = nullptr;
            if (!trc_dev) { cudaMalloc(&trc_dev, 8 * 32 * 12 * sizeof(long long)); }
            cudaMemsetAsync(trc_dev, 0, 8 * 32 * 12 * sizeof(long long), st);
            nostate = trc_dev;
#endif
            void* args[] = {(void*)&xi, (void*)&sk, (void*)&xo, (void*)&h->hp->dec[i], (void*)&Tl, (void*)&seg_i, (void*)&Bi, (void*)&nostate, (void*)&blk0, (void*)&n0};
            cudaLaunchKernel(tc_dec_fn(nwg), g, dim3(128 * nwg), args, gtt_smem_floats(nwg) * sizeof(float), st);
            h->launches++;
#ifdef GTCRNB_DEC_TRACE
            if (i == 2 && getenv("GTCRNB_DEC_TRACE_PRINT")) {
                std::vector<long long> hbuf(8 * 32 * 12);
                cudaStreamSynchronize(st);
                cudaMemcpy(hbuf.data(), trc_dev, hbuf.size() * sizeof(long long), cudaMemcpyDeviceToHost);
                static const char* const names[12] = {"top", "loads", "pc1+split", "wait_st", "bar1", "issue", "gate", "mma_wait", "fetch+xchg", "bar2", "store", "epilogue"};
                for (int row = 0; row < 8; ++row) {
                    double acc[12] = {}; int n = 0;
                    for (int k = 0; k + 1 < 32; ++k) {
                        const long long* a = &hbuf[(row * 32 + k) * 12]; const long long* b = &hbuf[(row * 32 + k + 1) * 12];
                        if (!a[0] || !b[0]) continue;
                        for (int j = 0; j < 11; ++j) acc[j + 1] += (double)(a[j + 1] - a[j]);
                        acc[0] += (double)(b[0] - a[11]);
                        ++n;
                    }
                    if (!n) continue;
                    double tot = 0; for (double v : acc) tot += v;
                    fprintf(stderr, "wg %d thread %s: step %.0f cycles:", row / 2, row & 1 ? "row40" : "issuer", tot / n);
                    for (int j = 0; j < 12; ++j) fprintf(stderr, " %s %.0f", j == 0 ? "loop" : names[j], acc[j] / n);
                    fprintf(stderr, "\n");
                }
            }
#endif
        }
    } else {
        const GtVariant& v = kGtVariants[h->gt_dec];
        const int seg = pick_seg_len(B, Ti, v.ps, GT_WARMUP, h->sm_count * h->occ_gt_dec);
        const dim3 g((unsigned)((Ti + seg - 1) / seg), (unsigned)B);
        for (int i = 0; i < 3; ++i) {
            // de_i consumes (previous + en_outs[4 - i]) : skips en4, en3, en2 = act[3], act[2], act[1]
            ProfScope ps(h, KC_DEC_GT, nframes, st);
            launch_gtconv<true>(v, g, st, h->act[5 + i], h->act[3 - i], h->act[6 + i], h->hp->dec[i], Ti, seg, (int)B);
            h->launches++;
        }
    }
}

// Fused wav -> wav mode (wav != nullptr): the STFT runs inside k_front (spectrum parked in h->spec_in for the mask stage)
// and the iSTFT + overlap-add inside k_back (wav_out must be zeroed by the caller).  `phases` / [u0, u0 + nu) select a part of the
// model (host pipeline); the body always covers the B utterances of the workspace.
int run_model(gtcrnb_handle* h, const float* spec, const int64_t is[4], int64_t B, int64_t T, float* out,
              const int64_t os[4], cudaStream_t st, const float* wav = nullptr, int64_t wav_stride = 0, int L = 0,
              float* wav_out = nullptr, int64_t wav_out_stride = 0, int phases = PH_ALL, int64_t u0 = 0, int64_t nu = -1) {
    if (nu < 0) nu = B - u0;
    const bool fused = wav != nullptr || wav_out != nullptr;
    if (phases & PH_FRONT) {
        if (h->variant == 1) launch_front<1>(h, *h->hpf, spec, is, T, u0, nu, st, wav, wav_stride, L);
        else launch_front<0>(h, *h->hp, spec, is, T, u0, nu, st, wav, wav_stride, L);
    }
    if (phases & PH_BODY) {
        if (h->variant == 1) {
            static const int kCls[4] = {KC_FULL_ENC_GT, KC_DP_INTRA, KC_DP_INTER, KC_FULL_DEC_GT};
            h->launches += gf::run_body(h->blob, *h->hpf, h->act, B, (int)T, nullptr, st, [&](int c, int64_t fr) { return ProfScope(h, kCls[c], fr, st); });
        } else
            launch_body_micro(h, B, T, st);
    }
    if (phases & PH_BACK) {
        if (h->variant == 1) launch_back<1>(h, *h->hpf, spec, is, T, u0, nu, out, os, st, fused, wav_out, wav_out_stride);
        else launch_back<0>(h, *h->hp, spec, is, T, u0, nu, out, os, st, fused, wav_out, wav_out_stride);
    }
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

gs::Ws stream_ws(gtcrnb_handle* h) {
    gs::Ws ws;
    ws.en0 = h->en0;
    for (int i = 0; i < kNumAct33; ++i) ws.act[i] = h->act[i];
    ws.spec_in = h->spec_in;
    ws.spec_out = h->spec_out;
    return ws;
}

int64_t chunk_utts(int64_t B, int64_t T) { return std::max<int64_t>(1, std::min<int64_t>({B, kMaxChunkFrames / T, (int64_t)65535})); }

}  // namespace

extern "C" {

int gtcrnb_abi_version(void) { return kAbiVersion; }
size_t gtcrnb_blob_floats(void) { return (size_t)blob::kBlobFloats; }
const char* gtcrnb_last_error(void) { return g_err.c_str(); }
int64_t gtcrnb_launch_count(const gtcrnb_handle* h) { return h ? h->launches : 0; }
int gtcrnb_tc_mask(const gtcrnb_handle* h) { return (h && h->variant == 0) ? h->tc_mask : 0; }

void* gtcrnb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void* gtcrnb_host_alloc_wc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocWriteCombined) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void gtcrnb_host_free(void* p) { if (p) cudaFreeHost(p); }

size_t gtcrnb_blob_floats_variant(int variant) { return variant == 1 ? (size_t)blobf::kBlobFloats : variant == 0 ? (size_t)blob::kBlobFloats : 0; }

int gtcrnb_create(const float* blob_host, size_t n_floats, gtcrnb_handle** out) { return gtcrnb_create_variant(blob_host, n_floats, GTCRNB_VARIANT_MICRO, out); }

int gtcrnb_create_variant(const float* blob_host, size_t n_floats, int variant, gtcrnb_handle** out) {
    if (!blob_host || !out) return fail(GTCRNB_E_INVALID, "gtcrnb_create: NULL argument");
    if (variant != GTCRNB_VARIANT_MICRO && variant != GTCRNB_VARIANT_FULL) return fail(GTCRNB_E_INVALID, "gtcrnb_create: unknown variant %d", variant);
    const int want_n = variant == 1 ? blobf::kBlobFloats : blob::kBlobFloats;
    const float want_v = variant == 1 ? blobf::kVersion : blob::kVersion;
    if (n_floats != (size_t)want_n) return fail(GTCRNB_E_BLOB, "weight blob has %zu floats, expected %d (variant %d)", n_floats, want_n, variant);
    if (blob_host[0] != want_v || blob_host[1] != (float)want_n)
        return fail(GTCRNB_E_BLOB, "weight blob version %.1f / size %.0f does not match library (%.1f / %d, variant %d)", blob_host[0], blob_host[1], want_v, want_n, variant);
    int dev = 0;
    CU(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(GTCRNB_E_ARCH, "device %s is sm_%d%d; this library contains sm_100a code only", prop.name, prop.major, prop.minor);
    gtcrnb_handle* h = new gtcrnb_handle();
    h->variant = variant;
    h->device = dev;
    h->sm_count = prop.multiProcessorCount;
    if (cudaMalloc(&h->blob, sizeof(float) * n_floats) != cudaSuccess) { delete h; return fail(GTCRNB_E_NOMEM, "cudaMalloc(blob) failed"); }
    if (cudaMemcpy(h->blob, blob_host, sizeof(float) * n_floats, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(h->blob); delete h; return fail(GTCRNB_E_CUDA, "cudaMemcpy(blob) failed");
    }
    using namespace gk;
    struct { const void* fn; int floats; } attrs[] = {
        {(const void*)k_front<false, 0>, fr_smem_floats(false)}, {(const void*)k_front<true, 0>, fr_smem_floats(true)},
        {(const void*)k_front<false, 0, true>, fr_smem_floats(false, true)}, {(const void*)k_front<true, 0, true>, fr_smem_floats(true, true)},
        {(const void*)k_gtcn<2, false>, wc_smem_floats(2)}, {(const void*)k_gtcn<2, true>, wc_smem_floats(2, true)},
        {(const void*)k_back<false, 0>, bk_smem_floats(false)}, {(const void*)k_back<true, 0>, bk_smem_floats(true)},
        {(const void*)k_back<false, 0, true>, bk_smem_floats(false, true)}, {(const void*)k_back<true, 0, true>, bk_smem_floats(true, true)},
        {(const void*)k_stft, 0},
        {(const void*)k_istft, 0},
        {tc_gtcn_fn(1, false), tct_smem_floats(1)}, {tc_gtcn_fn(2, false), tct_smem_floats(2)}, {tc_gtcn_fn(4, false), tct_smem_floats(4)},
        {tc_dec_fn(1), gtt_smem_floats(1)}, {tc_dec_fn(2), gtt_smem_floats(2)}, {tc_dec_fn(4), gtt_smem_floats(4)}, {tc_dec_fn(5), gtt_smem_floats(5)},
        {(const void*)k_gtconv_dec_tc<4, true>, gtt_smem_floats(4)},
        {tc_gtcn_fn(1, true), tct_smem_floats(1)}, {tc_gtcn_fn(2, true), tct_smem_floats(2)}, {tc_gtcn_fn(4, true), tct_smem_floats(4)},
    };
    std::vector<std::pair<const void*, int>> all;
    for (auto& a : attrs) all.push_back({a.fn, a.floats});
    for (const GtVariant& v : kGtVariants) { all.push_back({v.enc, v.smem_enc / 4}); all.push_back({v.dec, v.smem_dec / 4}); }
    for (auto& pr : all) {
        struct { const void* fn; int floats; } a{pr.first, pr.second};
        cudaError_t e = cudaFuncSetAttribute(a.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, a.floats * (int)sizeof(float));
        // without this the driver picks the smallest carveout that fits ONE CTA and occupancy collapses (ncu: occ_smem = 1)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(a.fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { gtcrnb_destroy(h); return fail(GTCRNB_E_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); }
    }
    if (const char* e = getenv("GTCRNB_TC_DEC_CARVEOUT")) {        // percent of the maximum shared-memory carve-out for the decoder TC kernel (rest = L1)
        const int pct = std::max(0, std::min(100, atoi(e)));
        for (int nwg : {1, 2, 4, 5}) cudaFuncSetAttribute(tc_dec_fn(nwg), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    if (const char* e = getenv("GTCRNB_GT_ENC_CARVEOUT")) {
        const int pct = std::max(0, std::min(100, atoi(e)));
        for (const GtVariant& v : kGtVariants) cudaFuncSetAttribute(v.enc, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    if (const char* e = getenv("GTCRNB_TC_GTCN_CARVEOUT")) {
        const int pct = std::max(0, std::min(100, atoi(e)));
        for (int nwg : {1, 2, 4}) cudaFuncSetAttribute(tc_gtcn_fn(nwg, false), cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    int e2 = gs::stream_init_attributes();
    if (e2 != 0) { gtcrnb_destroy(h); return fail(GTCRNB_E_CUDA, "stream kernel attribute setup failed: %s", cudaGetErrorString((cudaError_t)e2)); }
    if (variant == 1) {
        const int e3 = gf::init_attributes();
        if (e3 != 0) { gtcrnb_destroy(h); return fail(GTCRNB_E_CUDA, "variant-1 kernel attribute setup failed: %s", cudaGetErrorString((cudaError_t)e3)); }
        h->hpf = new gf::HostParams();
        h->hpf->load(blob_host);
        *out = h;
        return GTCRNB_OK;
    }
    h->hp = new gs::HostParams();
    h->hp->load(blob_host);
    auto pick = [](const char* env, int dflt) {
        const char* e = getenv(env);
        if (!e) return dflt;
        const int v = atoi(e);                                  // e.g. 4 = (PS 4, constant bank), 14 = (PS 4, shared-memory weights)
        if (v >= 100) return std::min(kNumGtVariants - 1, v - 100);                      // 100 + i: variant index i (e.g. 107 = PS 2, G 4)
        for (int i = 0; i < kNumGtVariants; ++i) if (kGtVariants[i].ps == v % 10 && kGtVariants[i].smem_w == v / 10) return i;
        return dflt;
    };
    h->gt_enc = pick("GTCRNB_GT_PS_ENC", 7);
    h->gt_dec = pick("GTCRNB_GT_PS_DEC", 2);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_gt_enc, kGtVariants[h->gt_enc].enc, kGtVariants[h->gt_enc].threads, kGtVariants[h->gt_enc].smem_enc);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_gt_dec, kGtVariants[h->gt_dec].dec, kGtVariants[h->gt_dec].threads, kGtVariants[h->gt_dec].smem_dec);
    if (const char* e = getenv("GTCRNB_FUSE_FFT")) h->fuse_fft = atoi(e) != 0;
    if (const char* e = getenv("GTCRNB_TCN_SMEMW")) h->tcn_smem_w = atoi(e) != 0;
    if (const char* e = getenv("GTCRNB_STREAM_GRAPH")) h->stream_graphs = atoi(e) != 0;
    if (const char* e = getenv("GTCRNB_TC")) h->tc_mask = atoi(e);
    if (const char* e = getenv("GTCRNB_SEG_WAVES")) h->seg_waves = std::max(1, std::min(64, atoi(e)));
    if (h->tc_mask & 12) {
        if (cudaMalloc(&h->tcw, sizeof(uint32_t) * gk::TCW_WORDS) != cudaSuccess) { cudaGetLastError(); gtcrnb_destroy(h); return fail(GTCRNB_E_NOMEM, "cudaMalloc(tcw) failed"); }
        gk::k_prep_ends_tc<<<1, 256>>>(h->blob, h->tcw);
        if (cudaDeviceSynchronize() != cudaSuccess) { gtcrnb_destroy(h); return fail(GTCRNB_E_CUDA, "k_prep_ends_tc failed: %s", cudaGetErrorString(cudaGetLastError())); }
    }
    h->hp->tc_ends = (h->tc_mask & 4) != 0;
    h->hp->tc_back = (h->tc_mask & 8) != 0;
    if (const char* e = getenv("GTCRNB_STREAM_FUSE_ENC")) h->hp->fuse_enc = atoi(e) != 0;
    h->hp->tcw = h->tcw;
    // streaming: the decoder blocks of a hop follow the offline choice (bit 1) unless GTCRNB_TC_STREAM=0
    h->hp->tc_stream = (h->tc_mask & 2) != 0;
    if (const char* e = getenv("GTCRNB_TC_STREAM")) h->hp->tc_stream = atoi(e) != 0;
    h->hp->sm_count = h->sm_count;
    if (const char* e = getenv("GTCRNB_TC_DEC_NWG")) { const int v = atoi(e); h->tc_dec_nwg = v >= 5 ? 5 : v >= 4 ? 4 : v >= 2 ? 2 : 1; }
    if (const char* e = getenv("GTCRNB_TC_GTCN_NWG")) { const int v = atoi(e); h->tc_gtcn_nwg = v >= 4 ? 4 : v >= 2 ? 2 : 1; }
    {
        const int nwg = h->tc_gtcn_nwg;
        if (const char* e = getenv("GTCRNB_TC_STACK")) h->tc_stack = atoi(e) != 0;
        const void* fn = tc_gtcn_fn(nwg, h->tc_stack);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_tcn_tc, fn, 128 * nwg, tct_smem_floats(nwg) * sizeof(float));
        h->occ_tcn_tc = std::max(1, std::min(h->occ_tcn_tc, 8 / nwg));       // 64 TMEM columns per warpgroup, 512 per SM
    }
    if (h->tcn_smem_w) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_tcn, k_gtcn<2, true>, WC_WARPS * 32, wc_smem_floats(2, true) * sizeof(float));
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->occ_tcn, k_gtcn<2, false>, WC_WARPS * 32, wc_smem_floats(2) * sizeof(float));
    h->occ_gt_enc = std::max(1, h->occ_gt_enc); h->occ_gt_dec = std::max(1, h->occ_gt_dec); h->occ_tcn = std::max(1, h->occ_tcn);
    *out = h;
    return GTCRNB_OK;
}

void gtcrnb_destroy(gtcrnb_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    drop_step_graphs(h);
    if (h->s_cap) cudaStreamDestroy(h->s_cap);
    auto freep = [](float* p) { if (p) cudaFree(p); };
    select_workspace(h, 0);
    freep(h->blob); freep(h->en0); freep(h->spec_in); freep(h->spec_out);
    if (h->tcw) cudaFree(h->tcw);
    for (auto p : h->act) freep(p);
    for (int w = 1; w < gtcrnb_handle::kNumWs; ++w) {
        freep(h->sets[w].en0); freep(h->sets[w].spec_in); freep(h->sets[w].spec_out);
        for (auto p : h->sets[w].act) freep(p);
    }
    delete h->hp;
    delete h->hpf;
    for (int i = 0; i < gtcrnb_handle::kStage; ++i) {
        freep(h->stage_in[i]); freep(h->stage_out[i]);
        if (h->pcm_in[i]) cudaFree(h->pcm_in[i]);
        if (h->pcm_out[i]) cudaFree(h->pcm_out[i]);
        for (cudaEvent_t e : {h->ev_fronts[i], h->ev_body[i], h->ev_backs[i], h->ev_d2h[i]}) if (e) cudaEventDestroy(e);
    }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    for (cudaStream_t st : {h->s_h2d, h->s_front, h->s_body, h->s_body2, h->s_back, h->s_d2h}) if (st) cudaStreamDestroy(st);
    if (h->ev_user) cudaEventDestroy(h->ev_user);
    delete h;
}

int gtcrnb_forward_spec(gtcrnb_handle* h, const float* spec_dev, const int64_t in_strides[4], int64_t B, int64_t T,
                        float* out_dev, const int64_t out_strides[4], void* stream) {
    if (B < 0 || T < 0 || T > (1 << 24)) return fail(GTCRNB_E_INVALID, "gtcrnb_forward_spec: bad shape B=%lld T=%lld", (long long)B, (long long)T);
    if (h && (B == 0 || T == 0)) return GTCRNB_OK;           // empty batch: nothing to do (pointers may be NULL)
    if (!h || !spec_dev || !out_dev || !in_strides || !out_strides) return fail(GTCRNB_E_INVALID, "gtcrnb_forward_spec: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Bc = chunk_utts(B, T);
    int rc = ensure_workspace(h, Bc * T);
    if (rc) return rc;
    h->last_B = Bc; h->last_T = T;
    for (int64_t b0 = B - (B % Bc ? B % Bc : Bc); b0 >= 0; b0 -= Bc) {   // last chunk first so chunk 0 stays in the workspace for debug taps
        const int64_t nb = std::min(Bc, B - b0);
        rc = run_model(h, spec_dev + b0 * in_strides[0], in_strides, nb, T, out_dev + b0 * out_strides[0], out_strides, st);
        if (rc) return rc;
    }
    note_user_work(h, st);
    return GTCRNB_OK;
}

int gtcrnb_stft(gtcrnb_handle* h, const float* wav_dev, int64_t B, int64_t L, int64_t in_stride, float* spec_dev, void* stream) {
    if (!h || !wav_dev || !spec_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_stft: NULL argument");
    if (L <= 256 || L > (int64_t)1 << 30 || B < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stft: need 256 < L < 2^30 (reflect padding), got L=%lld", (long long)L);
    if (B == 0) return GTCRNB_OK;
    const int T = (int)(1 + L / gk::kHop);
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
        const int64_t nb = std::min<int64_t>(65535, B - b0);
        ProfScope ps(h, KC_STFT, nb * T, (cudaStream_t)stream);
        gk::k_stft<<<dim3((T + gk::STFT_FPB - 1) / gk::STFT_FPB, (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(
            wav_dev + b0 * in_stride, in_stride, (int)L, T, h->blob, spec_dev + b0 * (int64_t)T * gk::kSpecFrame);
        h->launches++;
    }
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

int gtcrnb_istft(gtcrnb_handle* h, const float* spec_dev, int64_t B, int64_t T, float* wav_dev, int64_t out_stride, void* stream) {
    if (!h || !spec_dev || !wav_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_istft: NULL argument");
    if (T < 1 || B < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_istft: bad shape");
    if (B == 0 || T == 1) return GTCRNB_OK;
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
        const int64_t nb = std::min<int64_t>(65535, B - b0);
        ProfScope ps(h, KC_ISTFT, nb * T, (cudaStream_t)stream);
        gk::k_istft<<<dim3((unsigned)((T - 1 + gk::ISTFT_HOPS - 1) / gk::ISTFT_HOPS), (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(
            spec_dev + b0 * T * gk::kSpecFrame, T * gk::kSpecFrame, 2, gk::kSpecFrame, 1, (int)T, h->blob, wav_dev + b0 * out_stride, out_stride);
        h->launches++;
    }
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

int gtcrnb_enhance(gtcrnb_handle* h, const float* wav_dev, int64_t B, int64_t L, int64_t in_stride, float* out_dev,
                   int64_t out_stride, void* stream) {
    if (L <= 256 || L > (int64_t)1 << 30 || B < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_enhance: need 256 < L < 2^30 (reflect padding), got L=%lld", (long long)L);
    if (h && B == 0) return GTCRNB_OK;
    if (!h || !wav_dev || !out_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_enhance: NULL argument");
    const int64_t T = 1 + L / gk::kHop;
    const int64_t Bc = chunk_utts(B, T);
    int rc = ensure_workspace(h, Bc * T);
    if (rc) return rc;
    h->last_B = Bc; h->last_T = T;
    const int64_t is[4] = {T * gk::kSpecFrame, 2, gk::kSpecFrame, 1};
    const int64_t Lout = gk::kHop * (T - 1);
    // fused STFT / iSTFT path needs 8-byte aligned output rows (vector RED.ADD); otherwise use the separate kernels
    const bool fuse = h->fuse_fft && (out_stride % 2 == 0) && ((reinterpret_cast<uintptr_t>(out_dev) & 7) == 0);
    for (int64_t b0 = B - (B % Bc ? B % Bc : Bc); b0 >= 0; b0 -= Bc) {
        const int64_t nb = std::min(Bc, B - b0);
        if (fuse) {
            float* o = out_dev + b0 * out_stride;
            if (out_stride == Lout) CU(cudaMemsetAsync(o, 0, sizeof(float) * nb * Lout, (cudaStream_t)stream));
            else CU(cudaMemset2DAsync(o, sizeof(float) * out_stride, 0, sizeof(float) * Lout, nb, (cudaStream_t)stream));
            rc = run_model(h, nullptr, is, nb, T, nullptr, is, (cudaStream_t)stream, wav_dev + b0 * in_stride, in_stride, (int)L, o, out_stride);
            if (rc) return rc;
            continue;
        }
        rc = gtcrnb_stft(h, wav_dev + b0 * in_stride, nb, L, in_stride, h->spec_in, stream);
        if (rc) return rc;
        rc = run_model(h, h->spec_in, is, nb, T, h->spec_out, is, (cudaStream_t)stream);
        if (rc) return rc;
        rc = gtcrnb_istft(h, h->spec_out, nb, T, out_dev + b0 * out_stride, out_stride, stream);
        if (rc) return rc;
    }
#ifdef GTCRNB_FB_TRACE
    if (getenv("GTCRNB_FB_TRACE_PRINT")) {
        cudaStreamSynchronize((cudaStream_t)stream);
        long long t[2][3][16];
        cudaMemcpyFromSymbol(t, gk::g_fb_trace, sizeof(t));
        static const char* const fn[15] = {"start", "loads issued", "bar(tables)", "FFT+feat", "bar", "pads", "bar", "ERB", "bar", "SFE", "bar", "en0", "bar", "en1 mma+epilogue", "bar"};
        static const char* const bn[11] = {"start", "phase a (loads, x planes)", "bar", "de3", "bar", "de4", "bar", "mask", "iFFT prep", "iFFT", "OLA"};
        for (int k = 0; k < 2; ++k)
            for (int w = 0; w < 3; ++w) {
                fprintf(stderr, "%s warp %d:", k ? "k_back" : "k_front", 4 * w);
                const int n = k ? 11 : 15;
                for (int i = 1; i < n; ++i) fprintf(stderr, " %s %lld |", k ? bn[i] : fn[i], t[k][w][i] && t[k][w][i - 1] ? t[k][w][i] - t[k][w][i - 1] : -1);
                fprintf(stderr, " total %lld\n", t[k][w][n - 1] - t[k][w][0]);
            }
        long long g[4][16];
        cudaMemcpyFromSymbol(g, gk::g_tcn_trace, sizeof(g));
        static const char* const gn[8] = {"", "conv1: split, st, bar, MMAs, wait", "ld + bias + PReLU", "Y store", "depthwise taps", "conv3: split, st, bar, MMAs, wait", "history", "ld + residual + PReLU"};
        for (int w = 0; w < 4; ++w) {
            fprintf(stderr, "k_gtcn_tc warp %d (one TCN of a 32-frame tile):", w);
            for (int i = 1; i < 8; ++i) fprintf(stderr, " %s %lld |", gn[i], g[w][i] - g[w][i - 1]);
            fprintf(stderr, " total %lld\n", g[w][7] - g[w][0]);
        }
    }
#endif
    note_user_work(h, (cudaStream_t)stream);
    return GTCRNB_OK;
}

}  // extern "C" (reopened below): the shared host pipeline is a C++ function

namespace {
// H2D / compute / D2H pipeline behind gtcrnb_enhance_host (float32 samples) and gtcrnb_enhance_host_pcm16 (PCM16 samples on the
// wire, converted on the device: x / 32768 in, floor(y * 32768) clamped out -- half the PCIe bytes of the float contract).
//
// Two granularities (round 2).  The per-frame ends of the model (front: STFT .. en1; back: de3 .. mask, iSTFT) are frame-parallel
// and run efficiently on a few utterances, while the walking kernels in between want hundreds of utterances per launch (one CTA
// or warp per utterance column; small batches are split into time segments that each pay a warm-up: 192 utterances per launch
// cost 10 % more per utterance than 1024, 48 cost 35 % more, profiles/batch_eff_r2.txt).  So the batch is cut into GROUPS
// (sized by the schedule below, one activation workspace + staging pair each, kStage in flight) and every group into COPY CHUNKS:
//   s_h2d   : H2D of copy chunk c                          s_front : front(c) as soon as its copy has landed
//   s_body  : walking kernels of the whole group once its last front is done
//   s_back  : back(c) per copy chunk                        s_d2h   : D2H of copy chunk c as soon as back(c) is done
// The first H2D and the last D2H cannot overlap with anything, so groups ramp up from g1 and down to glast utterances.
struct HostSchedule { int64_t copy; std::vector<int64_t> groups; int64_t gmax; };
HostSchedule host_schedule(int64_t B, int64_t L, int variant) {
    auto envi = [](const char* name, int dflt, int lo, int hi) { const char* e = getenv(name); return e ? std::max(lo, std::min(hi, atoi(e))) : dflt; };
    auto units = [&](int64_t u) { return std::max<int64_t>(1, (u * 160000) / L); };        // knobs are in 10-second-utterance equivalents
    HostSchedule s;
    s.copy = units(envi("GTCRNB_HOST_COPY", 48, 1, 4096));
    const int64_t gmax = std::min<int64_t>({B, units(envi("GTCRNB_HOST_CHUNK", 384, 1, 65535)), std::max<int64_t>(1, kMaxChunkFrames / (1 + L / gk::kHop))});
    // variant 1's GTConv / DPGRNN kernels are persistent over time (one CTA or warp per utterance): small groups leave SMs empty
    const int64_t g1 = std::min(gmax, units(envi("GTCRNB_HOST_G1", variant == 1 ? 96 : 48, 1, 65535)));
    const int64_t gl = std::min(gmax, units(envi("GTCRNB_HOST_GLAST", variant == 1 ? 96 : 24, 1, 65535)));
    const double up = envi("GTCRNB_HOST_GROWTH", 200, 101, 1000) / 100.0, down = envi("GTCRNB_HOST_DECAY", 160, 101, 1000) / 100.0;
    std::vector<int64_t> head, tail;
    int64_t used = 0;
    double a = (double)g1, b = (double)gl;
    // grow both ramps alternately while one more pair still leaves a full group for the middle
    for (bool more = envi("GTCRNB_HOST_RAMP", 1, 0, 1) != 0; more;) {
        more = false;
        if ((int64_t)a < gmax && used + (int64_t)a + gmax <= B) { head.push_back((int64_t)a); used += (int64_t)a; a *= up; more = true; }
        if ((int64_t)b < gmax && used + (int64_t)b + gmax <= B) { tail.push_back((int64_t)b); used += (int64_t)b; b *= down; more = true; }
    }
    s.groups = head;
    const int64_t left = B - used, nmid = (left + gmax - 1) / gmax;                       // the middle in equal groups of at most gmax
    for (int64_t i = 0; i < nmid; ++i) s.groups.push_back(left / nmid + (i < left % nmid ? 1 : 0));
    for (auto it = tail.rbegin(); it != tail.rend(); ++it) s.groups.push_back(*it);
    s.gmax = 0;
    for (int64_t g : s.groups) s.gmax = std::max(s.gmax, g);
    return s;
}

int host_pipeline(gtcrnb_handle* h, const void* wav_host, int64_t B, int64_t L, void* out_host, bool pcm) {
    if (L <= 256 || L > (int64_t)1 << 30 || B < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_enhance_host: need 256 < L < 2^30");
    if (h && B == 0) return GTCRNB_OK;                            // empty batch: torch hands out NULL data pointers
    if (!h || !wav_host || !out_host) return fail(GTCRNB_E_INVALID, "gtcrnb_enhance_host: NULL argument");
    constexpr int kStage = gtcrnb_handle::kStage;
    const int64_t T = 1 + L / gk::kHop, Lout = gk::kHop * (T - 1);
    if (!h->s_h2d) {
        // the per-chunk front / back kernels feed the copy engines and the next group: they go ahead of the walking kernels
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        const bool prio = getenv("GTCRNB_HOST_PRIO") ? atoi(getenv("GTCRNB_HOST_PRIO")) != 0 : true;
        CU(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithPriority(&h->s_front, cudaStreamNonBlocking, prio ? hi : lo));
        CU(cudaStreamCreateWithPriority(&h->s_body, cudaStreamNonBlocking, lo));
        CU(cudaStreamCreateWithPriority(&h->s_body2, cudaStreamNonBlocking, lo));
        CU(cudaStreamCreateWithPriority(&h->s_back, cudaStreamNonBlocking, prio ? hi : lo));
        CU(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
        for (int i = 0; i < kStage; ++i)
            for (cudaEvent_t* e : {&h->ev_fronts[i], &h->ev_body[i], &h->ev_backs[i], &h->ev_d2h[i]}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    }
    const HostSchedule sch = host_schedule(B, L, h->variant);
    const size_t need_in = (size_t)sch.gmax * L, need_out = (size_t)sch.gmax * Lout;
    if (need_in > h->stage_in_floats || need_out > h->stage_out_floats) {
        for (int i = 0; i < kStage; ++i) {
            if (h->stage_in[i]) cudaFree(h->stage_in[i]);
            if (h->stage_out[i]) cudaFree(h->stage_out[i]);
            h->stage_in[i] = h->stage_out[i] = nullptr;
        }
        h->stage_in_floats = h->stage_out_floats = 0;
        for (int i = 0; i < kStage; ++i) {
            if (cudaMalloc(&h->stage_in[i], need_in * sizeof(float)) != cudaSuccess ||
                cudaMalloc(&h->stage_out[i], need_out * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                return fail(GTCRNB_E_NOMEM, "staging allocation failed");
            }
        }
        h->stage_in_floats = need_in; h->stage_out_floats = need_out;
    }
    if (pcm && (need_in > h->pcm_in_elems || need_out > h->pcm_out_elems)) {
        for (int i = 0; i < kStage; ++i) {
            if (h->pcm_in[i]) cudaFree(h->pcm_in[i]);
            if (h->pcm_out[i]) cudaFree(h->pcm_out[i]);
            h->pcm_in[i] = h->pcm_out[i] = nullptr;
        }
        h->pcm_in_elems = h->pcm_out_elems = 0;
        for (int i = 0; i < kStage; ++i) {
            if (cudaMalloc(&h->pcm_in[i], need_in * sizeof(int16_t)) != cudaSuccess ||
                cudaMalloc(&h->pcm_out[i], need_out * sizeof(int16_t)) != cudaSuccess) {
                cudaGetLastError();
                return fail(GTCRNB_E_NOMEM, "PCM16 staging allocation failed");
            }
        }
        h->pcm_in_elems = need_in; h->pcm_out_elems = need_out;
    }
    // every group slot owns an activation workspace; (re)allocation synchronises the device, so do it before anything is queued
    for (int w = 0; w < std::min<int>(kStage, (int)sch.groups.size()); ++w) {
        select_workspace(h, w);
        const int rc = ensure_workspace(h, sch.gmax * T);
        if (rc) { select_workspace(h, 0); return rc; }
    }
    const size_t esz = pcm ? sizeof(int16_t) : sizeof(float);
    // the fused STFT / iSTFT kernels are the pipeline's front / back; with GTCRNB_FUSE_FFT=0 a group is one copy chunk through gtcrnb_enhance
    const bool split = h->fuse_fft;
    if (h->user_pending) {                                        // work queued on a caller stream still owns workspace 0
        CU(cudaStreamWaitEvent(h->s_front, h->ev_user, 0));
        CU(cudaStreamWaitEvent(h->s_body, h->ev_user, 0));
        CU(cudaStreamWaitEvent(h->s_body2, h->ev_user, 0));
        h->user_pending = false;
    }
    struct PipelineFlag { gtcrnb_handle* h; PipelineFlag(gtcrnb_handle* h_) : h(h_) { h->in_pipeline = true; } ~PipelineFlag() { h->in_pipeline = false; select_workspace(h, 0); } } pflag(h);
    // a failed launch must not leave work of this call running on the internal streams behind the caller's back
    auto drain = [&](int rc_) { for (cudaStream_t st_ : {h->s_h2d, h->s_front, h->s_body, h->s_body2, h->s_back, h->s_d2h}) if (st_) cudaStreamSynchronize(st_); cudaGetLastError(); return rc_; };
    // GTCRNB_HOST_TRACE=1: per-chunk engine timeline on stderr (diagnostic; adds timing events, not for benchmarking)
    const bool trace = getenv("GTCRNB_HOST_TRACE") != nullptr;
    struct TraceRow { int group; int64_t utts; int kind; cudaEvent_t e0, e1; };    // kind 0 h2d, 1 front, 2 body, 3 back, 4 d2h
    std::vector<TraceRow> rows;
    auto tbegin = [&](int g, int64_t n, int kind, cudaStream_t st) { if (trace) { TraceRow r{g, n, kind, nullptr, nullptr}; cudaEventCreate(&r.e0); cudaEventCreate(&r.e1); cudaEventRecord(r.e0, st); rows.push_back(r); } };
    auto tend = [&](cudaStream_t st) { if (trace) cudaEventRecord(rows.back().e1, st); };
    size_t pool_next = 0;
    auto pool_event = [&]() -> cudaEvent_t {
        if (pool_next == h->ev_pool.size()) { cudaEvent_t e = nullptr; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); h->ev_pool.push_back(e); }
        return h->ev_pool[pool_next++];
    };
    int64_t b0 = 0;
    const int64_t is[4] = {T * gk::kSpecFrame, 2, gk::kSpecFrame, 1};
    const auto t_enq0 = std::chrono::steady_clock::now();
    // second walking-kernel stream (by group parity).  GTCRN-Micro: no gain, every walking kernel is sized to fill the SMs in one wave
    // (21.8 vs 22.3 ms); variant 1: its persistent one-CTA-per-utterance kernels leave most SMs idle on a 100-utterance group, two groups
    // side by side give 252 k against 140 k audio-s/s
    const bool two_body = getenv("GTCRNB_HOST_BODY2") ? atoi(getenv("GTCRNB_HOST_BODY2")) != 0 : h->variant == 1;
    for (size_t g = 0; g < sch.groups.size(); b0 += sch.groups[g], ++g) {
        const int slot = (int)(g % kStage);
        const int64_t ng = sch.groups[g];
        const int64_t cc = split ? sch.copy : ng;
        const bool reuse = g >= (size_t)kStage;
        select_workspace(h, slot);
        float* in_f = h->stage_in[slot];
        float* out_f = h->stage_out[slot];
        // ---- copies in, fronts
        if (reuse) {
            CU(cudaStreamWaitEvent(h->s_h2d, split ? h->ev_fronts[slot] : h->ev_body[slot], 0));   // input staging of group g - kStage consumed
            CU(cudaStreamWaitEvent(h->s_front, h->ev_backs[slot], 0));     // its workspace released by its backs
        }
        for (int64_t u0 = 0; u0 < ng; u0 += cc) {
            const int64_t nu = std::min(cc, ng - u0);
            tbegin((int)g, nu, 0, h->s_h2d);
            CU(cudaMemcpyAsync(pcm ? (void*)(h->pcm_in[slot] + u0 * L) : (void*)(in_f + u0 * L), (const char*)wav_host + (b0 + u0) * L * esz, esz * nu * L,
                               cudaMemcpyHostToDevice, h->s_h2d));
            tend(h->s_h2d);
            cudaEvent_t e = pool_event();
            CU(cudaEventRecord(e, h->s_h2d));
            CU(cudaStreamWaitEvent(h->s_front, e, 0));
            tbegin((int)g, nu, 1, h->s_front);
            if (pcm) {
                gk::k_pcm16_to_f32<<<h->sm_count * 8, 256, 0, h->s_front>>>(h->pcm_in[slot] + u0 * L, in_f + u0 * L, nu * L);
                h->launches++;
            }
            if (split) {
                const int rc = run_model(h, nullptr, is, ng, T, nullptr, is, h->s_front, in_f, L, (int)L, nullptr, 0, PH_FRONT, u0, nu);
                if (rc) return drain(rc);
            }
            tend(h->s_front);
        }
        CU(cudaEventRecord(h->ev_fronts[slot], h->s_front));
        // ---- walking kernels of the group (two streams by group parity: the partial waves of one group's kernels fill with the next group's)
        cudaStream_t sb = (two_body && (g & 1)) ? h->s_body2 : h->s_body;
        CU(cudaStreamWaitEvent(sb, h->ev_fronts[slot], 0));
        tbegin((int)g, ng, 2, sb);
        if (split) {
            const int rc = run_model(h, nullptr, is, ng, T, nullptr, is, sb, nullptr, 0, 0, nullptr, 0, PH_BODY);
            if (rc) return drain(rc);
        } else {
            if (reuse) CU(cudaStreamWaitEvent(sb, h->ev_d2h[slot], 0));
            const int rc = gtcrnb_enhance(h, in_f, ng, L, L, out_f, Lout, sb);
            if (rc) return drain(rc);
        }
        tend(sb);
        CU(cudaEventRecord(h->ev_body[slot], sb));
        // ---- backs, copies out
        CU(cudaStreamWaitEvent(h->s_back, h->ev_body[slot], 0));
        if (reuse) CU(cudaStreamWaitEvent(h->s_back, h->ev_d2h[slot], 0));  // output staging of group g - kStage drained
        if (split) CU(cudaMemsetAsync(out_f, 0, sizeof(float) * ng * Lout, h->s_back));           // the fused iSTFT overlap-adds onto zeros
        for (int64_t u0 = 0; u0 < ng; u0 += cc) {
            const int64_t nu = std::min(cc, ng - u0);
            tbegin((int)g, nu, 3, h->s_back);
            if (split) {
                const int rc = run_model(h, nullptr, is, ng, T, nullptr, is, h->s_back, nullptr, 0, (int)L, out_f, Lout, PH_BACK, u0, nu);
                if (rc) return drain(rc);
            }
            if (pcm) {
                gk::k_f32_to_pcm16<<<h->sm_count * 8, 256, 0, h->s_back>>>(out_f + u0 * Lout, h->pcm_out[slot] + u0 * Lout, nu * Lout);
                h->launches++;
            }
            tend(h->s_back);
            cudaEvent_t e = pool_event();
            CU(cudaEventRecord(e, h->s_back));
            CU(cudaStreamWaitEvent(h->s_d2h, e, 0));
            tbegin((int)g, nu, 4, h->s_d2h);
            CU(cudaMemcpyAsync((char*)out_host + (b0 + u0) * Lout * esz, pcm ? (const void*)(h->pcm_out[slot] + u0 * Lout) : (const void*)(out_f + u0 * Lout),
                               esz * nu * Lout, cudaMemcpyDeviceToHost, h->s_d2h));
            tend(h->s_d2h);
        }
        CU(cudaEventRecord(h->ev_backs[slot], h->s_back));
        CU(cudaEventRecord(h->ev_d2h[slot], h->s_d2h));
    }
    const double enq_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enq0).count();
    CU(cudaStreamSynchronize(h->s_d2h));
    CU(cudaStreamSynchronize(h->s_back));
    CU(cudaStreamSynchronize(h->s_body));
    CU(cudaStreamSynchronize(h->s_body2));
    CU(cudaStreamSynchronize(h->s_front));
    CU(cudaStreamSynchronize(h->s_h2d));
    if (trace) {
        static const char* const kKind[5] = {"h2d", "front", "body", "back", "d2h"};
        fprintf(stderr, "host enqueue of the whole schedule: %.3f ms\ngroups:", enq_ms);
        for (int64_t g : sch.groups) fprintf(stderr, " %lld", (long long)g);
        fprintf(stderr, "  copy chunk %lld\ngroup  utts  engine   start     end   (ms from the first H2D)\n", (long long)sch.copy);
        for (const auto& r : rows) {
            float t0 = 0.f, t1 = 0.f;
            cudaEventElapsedTime(&t0, rows[0].e0, r.e0); cudaEventElapsedTime(&t1, rows[0].e0, r.e1);
            fprintf(stderr, "%5d %5lld  %-6s %7.3f %7.3f\n", r.group, (long long)r.utts, kKind[r.kind], t0, t1);
        }
        for (const auto& r : rows) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    }
    return GTCRNB_OK;
}
}  // namespace

extern "C" {
int gtcrnb_enhance_host(gtcrnb_handle* h, const float* wav_host, int64_t B, int64_t L, float* out_host) {
    return host_pipeline(h, wav_host, B, L, out_host, false);
}
int gtcrnb_enhance_host_pcm16(gtcrnb_handle* h, const int16_t* wav_host, int64_t B, int64_t L, int16_t* out_host) {
    return host_pipeline(h, wav_host, B, L, out_host, true);
}

int gtcrnb_pcm16_to_f32(gtcrnb_handle* h, const int16_t* in_dev, float* out_dev, int64_t n, void* stream) {
    if (n < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_pcm16_to_f32: bad size");
    if (h && n == 0) return GTCRNB_OK;
    if (!h || !in_dev || !out_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_pcm16_to_f32: NULL argument");
    gk::k_pcm16_to_f32<<<h->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(in_dev, out_dev, n);
    h->launches++;
    CU(cudaGetLastError());
    return GTCRNB_OK;
}
int gtcrnb_f32_to_pcm16(gtcrnb_handle* h, const float* in_dev, int16_t* out_dev, int64_t n, void* stream) {
    if (n < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_f32_to_pcm16: bad size");
    if (h && n == 0) return GTCRNB_OK;
    if (!h || !in_dev || !out_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_f32_to_pcm16: NULL argument");
    gk::k_f32_to_pcm16<<<h->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(in_dev, out_dev, n);
    h->launches++;
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

int gtcrnb_debug_activation(gtcrnb_handle* h, const char* name, float* out_dev, size_t n_floats, void* stream) {
    if (!h || !name || !out_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_debug_activation: NULL argument");
    const float* src = nullptr;
    size_t per_frame = gk::kF2 * 16;
    if (!strcmp(name, "en0")) { src = h->en0; per_frame = gk::kF1 * 16; }
    static const char* const kActNamesFull[kNumAct33] = {"en1", "en2", "en3", "en4", "dp1", "dp2", "de0", "de1", "de2"};
    for (int i = 0; i < kNumAct33 && !src; ++i)
        if (!strcmp(name, h->variant == 1 ? kActNamesFull[i] : kActNames[i])) src = h->act[i];
    if (!src) return fail(GTCRNB_E_INVALID, "unknown activation '%s'", name);
    const size_t have = per_frame * (size_t)(h->last_B * h->last_T);
    if (n_floats > have) return fail(GTCRNB_E_INVALID, "activation '%s' holds %zu floats, %zu requested", name, have, n_floats);
    CU(cudaMemcpyAsync(out_dev, src, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GTCRNB_OK;
}

// ---------------------------------------------------------------- on-device quality metrics
int gtcrnb_pair_metrics(gtcrnb_handle* h, const float* ref_dev, int64_t ref_stride, const float* inf_dev, int64_t inf_stride,
                        int64_t B, int64_t L, float* out_dev, void* stream) {
    if (B < 0 || L <= 0) return fail(GTCRNB_E_INVALID, "gtcrnb_pair_metrics: bad shape");
    if (h && B == 0) return GTCRNB_OK;
    if (!h || !ref_dev || !inf_dev || !out_dev) return fail(GTCRNB_E_INVALID, "gtcrnb_pair_metrics: NULL argument");
    for (int64_t b0 = 0; b0 < B; b0 += 65535) {
        const int64_t nb = std::min<int64_t>(65535, B - b0);
        gk::k_pair_metrics<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(ref_dev + b0 * ref_stride, ref_stride, inf_dev + b0 * inf_stride,
                                                                           inf_stride, L, out_dev + 2 * b0);
        h->launches++;
    }
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

// ---------------------------------------------------------------- per-kernel profiling
int gtcrnb_profile_enable(gtcrnb_handle* h, int on) {
    if (!h) return fail(GTCRNB_E_INVALID, "gtcrnb_profile_enable: NULL handle");
    h->profiling = on != 0;
    return GTCRNB_OK;
}

int gtcrnb_profile_read(gtcrnb_handle* h, int reset, double* ms_out, int64_t* launches_out, int64_t* frames_out) {
    if (!h || !ms_out || !launches_out || !frames_out) return fail(GTCRNB_E_INVALID, "gtcrnb_profile_read: NULL argument");
    for (auto& ev : h->prof_events) {
        CU(cudaEventSynchronize(ev.e1));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        h->prof_ms[ev.cls] += ms; h->prof_launches[ev.cls] += 1; h->prof_frames[ev.cls] += ev.frames;
        cudaEventDestroy(ev.e0); cudaEventDestroy(ev.e1);
    }
    h->prof_events.clear();
    for (int i = 0; i < kNumKernelClasses; ++i) {
        ms_out[i] = h->prof_ms[i]; launches_out[i] = h->prof_launches[i]; frames_out[i] = h->prof_frames[i];
        if (reset) { h->prof_ms[i] = 0; h->prof_launches[i] = 0; h->prof_frames[i] = 0; }
    }
    return GTCRNB_OK;
}

int gtcrnb_profile_num_classes(void) { return kNumKernelClasses; }
const char* gtcrnb_profile_class_name(int i) {
    static const char* const names[kNumKernelClasses] = {"stft", "front", "enc_gtconv", "gtcn", "dec_gtconv", "back", "istft", "stream",
                                                         "full_enc_gtconv", "dp_intra", "dp_inter", "full_dec_gtconv"};
    return (i >= 0 && i < kNumKernelClasses) ? names[i] : "";
}

// ---------------------------------------------------------------- streaming
size_t gtcrnb_stream_state_bytes(int64_t N) { return N < 0 ? 0 : gs::state_bytes(N); }
size_t gtcrnb_stream_state_bytes_variant(int variant, int64_t N) { return N < 0 ? 0 : variant == 1 ? gf::state_bytes(N) : variant == 0 ? gs::state_bytes(N) : 0; }

int gtcrnb_stream_reset(gtcrnb_handle* h, void* state_dev, int64_t N, void* stream) {
    if (!h || !state_dev || N < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_reset: bad argument");
    CU(cudaMemsetAsync(state_dev, 0, h->variant == 1 ? gf::state_bytes(N) : gs::state_bytes(N), (cudaStream_t)stream));
    return GTCRNB_OK;
}

}  // extern "C"

namespace {
// kind 0: spectral frame in / out, kind 1: waveform hop in / out
int stream_step_impl(gtcrnb_handle* h, int kind, void* state, const float* in, float* out, int64_t N, cudaStream_t st);
int stream_step(gtcrnb_handle* h, int kind, void* state, const float* in, float* out, int64_t N, cudaStream_t st) {
    const int rc = stream_step_impl(h, kind, state, in, out, N, st);
    if (rc == GTCRNB_OK) note_user_work(h, st);
    return rc;
}
int stream_step_impl(gtcrnb_handle* h, int kind, void* state, const float* in, float* out, int64_t N, cudaStream_t st) {
    int rc = ensure_workspace(h, N);
    if (rc) return rc;
    auto direct = [&](cudaStream_t s) {
        if (h->variant == 1)
            return kind ? gf::step_wav(h->blob, *h->hpf, state, stream_ws(h), in, out, N, s) : gf::step_spec(h->blob, *h->hpf, state, stream_ws(h), in, out, N, s);
        return kind ? gs::step_wav(h->blob, *h->hp, state, stream_ws(h), in, out, N, s) : gs::step_spec(h->blob, *h->hp, state, stream_ws(h), in, out, N, s);
    };
    if (!h->stream_graphs || h->profiling || h->ws_cur != 0) {
        h->launches += direct(st);
        CU(cudaGetLastError());
        return GTCRNB_OK;
    }
    ++h->step_tick;
    for (auto& g : h->step_graphs)
        if (g.kind == kind && g.state == state && g.in == in && g.out == out && g.n == N) {
            g.used = h->step_tick;
            CU(cudaGraphLaunch(g.exec, st));
            h->launches += g.launches;
            return GTCRNB_OK;
        }
    // a hop is captured the second time its (state, in, out, N) shows up among the last few distinct ones: a caller that rotates over
    // several stream groups (StreamPool) gets one graph per group, pointers that change every hop never pay for a capture
    bool repeat = false;
    for (auto& l : h->seen_steps)
        if (l.kind == kind && l.state == state && l.in == in && l.out == out && l.n == N) { repeat = true; break; }
    if (!repeat) {
        if (h->seen_steps.size() >= 16) h->seen_steps.erase(h->seen_steps.begin());
        h->seen_steps.push_back({kind, state, in, out, N, nullptr, 0, 0});
        h->launches += direct(st);
        CU(cudaGetLastError());
        return GTCRNB_OK;
    }
    if (!h->s_cap) CU(cudaStreamCreateWithFlags(&h->s_cap, cudaStreamNonBlocking));
    CU(cudaStreamBeginCapture(h->s_cap, cudaStreamCaptureModeThreadLocal));
    const int n_launch = direct(h->s_cap);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->s_cap, &graph);
    cudaGraphExec_t exec = nullptr;
    if (e == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {                           // capture is an optimisation: fall back to plain launches
        cudaGetLastError();
        h->stream_graphs = false;
        h->launches += direct(st);
        CU(cudaGetLastError());
        return GTCRNB_OK;
    }
    if (h->step_graphs.size() >= 16) {                // small LRU
        size_t victim = 0;
        for (size_t i = 1; i < h->step_graphs.size(); ++i) if (h->step_graphs[i].used < h->step_graphs[victim].used) victim = i;
        cudaGraphExecDestroy(h->step_graphs[victim].exec);
        h->step_graphs.erase(h->step_graphs.begin() + victim);
    }
    h->step_graphs.push_back({kind, state, in, out, N, exec, n_launch, h->step_tick});
    CU(cudaGraphLaunch(exec, st));
    h->launches += n_launch;
    return GTCRNB_OK;
}
}  // namespace

extern "C" {
int gtcrnb_stream_step_spec(gtcrnb_handle* h, void* state_dev, const float* spec_dev, float* out_dev, int64_t N, void* stream) {
    if (!h || !state_dev || !spec_dev || !out_dev || N < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_step_spec: bad argument");
    if (N == 0) return GTCRNB_OK;
    return stream_step(h, 0, state_dev, spec_dev, out_dev, N, (cudaStream_t)stream);
}

int gtcrnb_stream_step_wav(gtcrnb_handle* h, void* state_dev, const float* hop_in_dev, float* hop_out_dev, int64_t N, void* stream) {
    if (!h || !state_dev || !hop_in_dev || !hop_out_dev || N < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_step_wav: bad argument");
    if (N == 0) return GTCRNB_OK;
    return stream_step(h, 1, state_dev, hop_in_dev, hop_out_dev, N, (cudaStream_t)stream);
}

int gtcrnb_stream_flush_wav(gtcrnb_handle* h, void* state_dev, const float* tail_dev, int64_t tail_len, float* out_dev, int64_t N, void* stream) {
    if (!h || !state_dev || !out_dev || N < 0 || tail_len < 0 || tail_len > 255 || (tail_len > 0 && !tail_dev))
        return fail(GTCRNB_E_INVALID, "gtcrnb_stream_flush_wav: bad argument (0 <= tail_len <= 255, tail_dev required when tail_len > 0)");
    if (N == 0) return GTCRNB_OK;
    if (N > 2147483647LL) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_flush_wav: too many streams");
    int rc = ensure_workspace(h, N);
    if (rc) return rc;
    if (h->variant == 1) h->launches += gf::flush_wav(h->blob, *h->hpf, state_dev, stream_ws(h), tail_dev, (int)tail_len, out_dev, N, (cudaStream_t)stream);
    else h->launches += gs::flush_wav(h->blob, *h->hp, state_dev, stream_ws(h), tail_dev, (int)tail_len, out_dev, N, (cudaStream_t)stream);
    CU(cudaGetLastError());
    note_user_work(h, (cudaStream_t)stream);
    return GTCRNB_OK;
}

int gtcrnb_stream_import_caches(gtcrnb_handle* h, void* state_dev, int64_t N, const float* conv_cache_dev, const float* tra_cache_dev,
                                const float* const tcn_cache_dev[8], void* stream) {
    if (!h || !state_dev || !conv_cache_dev || !tra_cache_dev || !tcn_cache_dev || N < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_import_caches: bad argument");
    if (h->variant != 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_import_caches: the reference cache layout exists for GTCRN-Micro only");
    if (N == 0) return GTCRNB_OK;
    h->launches += gs::convert_caches(true, state_dev, N, const_cast<float*>(conv_cache_dev), const_cast<float*>(tra_cache_dev),
                                      const_cast<float* const*>(tcn_cache_dev), (cudaStream_t)stream);
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

int gtcrnb_stream_export_caches(gtcrnb_handle* h, const void* state_dev, int64_t N, float* conv_cache_dev, float* tra_cache_dev,
                                float* const tcn_cache_dev[8], void* stream) {
    if (!h || !state_dev || !conv_cache_dev || !tra_cache_dev || !tcn_cache_dev || N < 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_export_caches: bad argument");
    if (h->variant != 0) return fail(GTCRNB_E_INVALID, "gtcrnb_stream_export_caches: the reference cache layout exists for GTCRN-Micro only");
    if (N == 0) return GTCRNB_OK;
    h->launches += gs::convert_caches(false, const_cast<void*>(state_dev), N, conv_cache_dev, tra_cache_dev, tcn_cache_dev, (cudaStream_t)stream);
    CU(cudaGetLastError());
    return GTCRNB_OK;
}

}  // extern "C"
