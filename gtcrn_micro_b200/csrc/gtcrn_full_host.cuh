// Host-side launch plumbing of the original-GTCRN-style variant (kernels: gtcrn_full.cuh, VAR = 1 front / back kernels).
#pragma once
#include <string.h>

#include "gtcrn_full.cuh"
#include "gtcrn_stream.cuh"

namespace gf {

constexpr gs::WavLayout kWavLayoutFull{kPerStream, kWavIn, kOla};

struct HostParams {
    gk::ParamBlock<blobf::FRONT_SIZE> front;
    gk::ParamBlock<blobf::GTF_SIZE> gt[6];
    gk::ParamBlock<blobf::DPA_SIZE> dpa[2];
    gk::ParamBlock<blobf::DPE_SIZE> dpe[2];
    gk::ParamBlock<blobf::BACK_SIZE> back;
    void load(const float* b) {
        memcpy(front.w, b + blobf::FRONT_OFF, sizeof(front.w));
        memcpy(back.w, b + blobf::BACK_OFF, sizeof(back.w));
        for (int i = 0; i < 6; ++i) memcpy(gt[i].w, b + blobf::GTF_OFF + i * blobf::GTF_SIZE, sizeof(gt[i].w));
        for (int i = 0; i < 2; ++i) {
            memcpy(dpa[i].w, b + blobf::DPA_OFF + i * blobf::DPA_SIZE, sizeof(dpa[i].w));
            memcpy(dpe[i].w, b + blobf::DPE_OFF + i * blobf::DPE_SIZE, sizeof(dpe[i].w));
        }
    }
};

// offline tiles: 2 frame slots per thread x 4 slot groups (8-frame tiles, 160 threads); streaming: one frame per call (64 threads)
constexpr int GTF_PS = 2, GTF_G = 4;

inline int init_attributes() {
    struct { const void* fn; int floats; } a[] = {
        {(const void*)kf_gtconv<false, GTF_PS, GTF_G>, gtf_smem_floats(GTF_PS, GTF_G)}, {(const void*)kf_gtconv<true, GTF_PS, GTF_G>, gtf_smem_floats(GTF_PS, GTF_G)},
        {(const void*)kf_gtconv<false, 1, 1>, gtf_smem_floats(1, 1)}, {(const void*)kf_gtconv<true, 1, 1>, gtf_smem_floats(1, 1)},
        {(const void*)kf_dp_intra, DPA_SMEM_FLOATS}, {(const void*)kf_dp_inter, DPE_SMEM_FLOATS},
        {(const void*)gk::k_front<false, 1>, gk::fr_smem_floats(false)}, {(const void*)gk::k_front<true, 1>, gk::fr_smem_floats(true)},
        {(const void*)gk::k_back<false, 1>, gk::bk_smem_floats(false)}, {(const void*)gk::k_back<true, 1>, gk::bk_smem_floats(true)},
    };
    for (auto& x : a) {
        cudaError_t e = cudaFuncSetAttribute(x.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, x.floats * (int)sizeof(float));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(x.fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

// The model between the front and back kernels: 3 encoder GTConv blocks, 2 DPGRNNs, 3 decoder GTConv blocks (8 + 2 launches).
// act[]: 0 en1, 1-3 en2-en4, 4 dp1, 5 dp2, 6-8 de0-de2; the intra halves of the DPGRNNs borrow act[6] / act[7] (free until the decoder).
// state == nullptr: offline, B utterances x T frames.  Otherwise one frame (T = 1) of B streams continuing `state`.
// `launch` is a callable (class, frames, fn) -> void so that the caller can bracket launches with profiling events.
template <class Launch>
inline int run_body(const float* blobp, const HostParams& hp, float* const act[9], int64_t B, int T, void* state, cudaStream_t st, Launch&& scope) {
    int launches = 0;
    const int64_t nframes = B * T;
    const bool stream = state != nullptr;
    for (int i = 0; i < 3; ++i) {
        auto sc = scope(0, nframes);
        if (stream) kf_gtconv<false, 1, 1><<<(unsigned)B, gtf_threads(1), gtf_smem_floats(1, 1) * sizeof(float), st>>>(act[i], nullptr, act[i + 1], hp.gt[i], T, gt_dil(i), state, i);
        else kf_gtconv<false, GTF_PS, GTF_G><<<(unsigned)B, gtf_threads(GTF_G), gtf_smem_floats(GTF_PS, GTF_G) * sizeof(float), st>>>(act[i], nullptr, act[i + 1], hp.gt[i], T, gt_dil(i), nullptr, i);
        ++launches;
    }
    for (int n = 0; n < 2; ++n) {
        const float* ln = blobp + blobf::DPLN_OFF + n * blobf::DPLN_SIZE;
        const float* xin = act[3 + n];
        {
            auto sc = scope(1, nframes);
            kf_dp_intra<<<(unsigned)((nframes + DPA_FR - 1) / DPA_FR), 128, DPA_SMEM_FLOATS * sizeof(float), st>>>(xin, act[6 + n], hp.dpa[n], ln + blobf::DPLN_A_W, nframes);
        }
        {
            auto sc = scope(2, nframes);
            kf_dp_inter<<<(unsigned)B, 128, DPE_SMEM_FLOATS * sizeof(float), st>>>(act[6 + n], act[4 + n], hp.dpe[n], ln + blobf::DPLN_E_W, T, state, n);
        }
        launches += 2;
    }
    for (int i = 0; i < 3; ++i) {
        auto sc = scope(3, nframes);
        if (stream) kf_gtconv<true, 1, 1><<<(unsigned)B, gtf_threads(1), gtf_smem_floats(1, 1) * sizeof(float), st>>>(act[5 + i], act[3 - i], act[6 + i], hp.gt[3 + i], T, gt_dil(3 + i), state, 3 + i);
        else kf_gtconv<true, GTF_PS, GTF_G><<<(unsigned)B, gtf_threads(GTF_G), gtf_smem_floats(GTF_PS, GTF_G) * sizeof(float), st>>>(act[5 + i], act[3 - i], act[6 + i], hp.gt[3 + i], T, gt_dil(3 + i), nullptr, 3 + i);
        ++launches;
    }
    return launches;
}

struct NoScope { int operator()(int, int64_t) const { return 0; } };

// one streaming hop on spectra already in `spec` [N][257][2]
inline int model_step(const float* blobp, const HostParams& hp, void* state, const gs::Ws& ws, const float* spec, float* out, int64_t N, cudaStream_t st) {
    using namespace gk;
    const unsigned gfr = (unsigned)((N + FR_NF - 1) / FR_NF), gbk = (unsigned)((N + BK_NF - 1) / BK_NF);
    k_front<false, 1><<<gfr, FR_THREADS, fr_smem_floats(false) * sizeof(float), st>>>(spec, kSpecFrame, 2, 0, 1, nullptr, 0, 0, nullptr, 1, N, blobp, hp.front, ws.en0, ws.act[0]);
    int launches = 1 + run_body(blobp, hp, ws.act, N, 1, state, st, NoScope{});
    k_back<false, 1><<<gbk, BK_THREADS, bk_smem_floats(false) * sizeof(float), st>>>(ws.act[8], ws.act[0], ws.en0, spec, kSpecFrame, 2, 0, 1,
                                                                                   out, kSpecFrame, 2, 0, 1, nullptr, 0, 1, N, blobp, hp.back);
    return launches + 1;
}
inline int step_spec(const float* blobp, const HostParams& hp, void* state, const gs::Ws& ws, const float* spec, float* out, int64_t N, cudaStream_t st) {
    const int launches = model_step(blobp, hp, state, ws, spec, out, N, st);
    gs::ks_advance<<<1, 1, 0, st>>>(state, 0);
    return launches + 1;
}
inline int step_wav(const float* blobp, const HostParams& hp, void* state, const gs::Ws& ws, const float* hop_in, float* hop_out, int64_t N, cudaStream_t st) {
    const unsigned g4 = (unsigned)((N + 3) / 4);
    gs::ks_stft_step<<<g4, 256, 0, st>>>(state, hop_in, blobp, ws.spec_in, N, kWavLayoutFull);
    const int launches = 1 + model_step(blobp, hp, state, ws, ws.spec_in, ws.spec_out, N, st);
    gs::ks_istft_step<<<g4, 256, 0, st>>>(state, ws.spec_out, blobp, hop_out, N, kWavLayoutFull);
    gs::ks_advance<<<1, 1, 0, st>>>(state, 1);
    return launches + 2;
}
inline int flush_wav(const float* blobp, const HostParams& hp, void* state, const gs::Ws& ws, const float* tail, int tail_len, float* out, int64_t N, cudaStream_t st) {
    float* v = out + N * 256;
    gs::ks_reflect_tail<<<(unsigned)N, 256, 0, st>>>(state, tail, tail_len, v, N, kWavLayoutFull);
    int launches = 1 + step_wav(blobp, hp, state, ws, v, out, N, st);
    launches += step_wav(blobp, hp, state, ws, v, v, N, st);
    return launches;
}

}  // namespace gf
