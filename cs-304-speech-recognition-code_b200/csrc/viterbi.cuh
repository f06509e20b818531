// Shared argument block of the Viterbi kernels.
#pragma once
#include "common.cuh"

namespace loe {

struct VitArgs {
    const float* scores; int ld;
    const int64_t* frm_off;
    const int32_t* tr_off; const int32_t* col; const float* band; const uint8_t* flags;
    const int32_t* utt_tr;
    int loop; float pen32; double pen64; int pen_f64;
    int8_t* path; float* end_scores; int max_ends; int32_t* best; float* best_score;
    uint8_t* bp_ws; int bp_in_smem; int max_frames; int max_pos;
    // optional fused label decoding (model_boundary.py:107-147); words == nullptr disables it
    const int32_t* word; const int32_t* word_lo; int skip_label; int8_t* words; int max_words; int32_t* count;
};

// One-warp-per-utterance kernel (viterbi_warp.cu).  Returns false when the utterances are too long
// for its shared-memory back-pointer store; the caller then uses the CTA-per-utterance kernel.
bool viterbi_warp_launch(const VitArgs& a, int n_utt, cudaStream_t s);

}  // namespace loe
