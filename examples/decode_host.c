/* Plain C caller of the host-buffer decoder (include/loe_b200.h): no Python, no torch.
 *
 *   gcc -O2 -Iinclude examples/decode_host.c -Lcs-304-speech-recognition-code_b200/lib -lloe_b200 \
 *       -Wl,-rpath,$PWD/cs-304-speech-recognition-code_b200/lib -o decode_host
 *   ./decode_host model_and_pcm.blob
 *
 * The blob is written by loe_speech_recognition._decoder.write_blob (model tables of a
 * HiddenMarkovModelInference + a PCM batch); the program prints one line of word ids per utterance.
 * tests/test_gpu_parity.py::test_c_program_decodes_like_python builds and runs it.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "loe_b200.h"

static void* slurp(FILE* f, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (!p || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short read\n"); exit(2); }
    return p;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s blob\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("blob"); return 2; }
    int64_t h[12];
    double penalty;
    if (fread(h, sizeof h, 1, f) != 1 || fread(&penalty, sizeof penalty, 1, f) != 1) { fprintf(stderr, "bad header\n"); return 2; }
    const int mel_na = (int)h[0], mel_nb = (int)h[1], n_states = (int)h[2], n_pos = (int)h[3], n_utt = (int)h[4],
              max_words = (int)h[5], pcm_format = (int)h[6], penalty_f64 = (int)h[7], skip_label = (int)h[8], n_chunks = (int)h[9];
    const int64_t n_samples = h[10];
    const int n_tiles = loe_emission_tc_tiles(n_states);
    int32_t* mel_bin = slurp(f, (size_t)(mel_na + mel_nb) * 32 * 4);
    float* mel_w = slurp(f, (size_t)(mel_na + mel_nb) * 32 * 4);
    float* b_packed = slurp(f, (size_t)n_tiles * 19200 * 4);
    float* cst_pad = slurp(f, (size_t)n_tiles * 6 * 4);
    int32_t* col = slurp(f, (size_t)n_pos * 4);
    float* band = slurp(f, (size_t)n_pos * 3 * 4);
    uint8_t* flags = slurp(f, (size_t)n_pos);
    int32_t* word = slurp(f, (size_t)n_pos * 4);
    int32_t* word_lo = slurp(f, (size_t)n_pos * 4);
    const size_t h16_bytes = (size_t)h[11];     /* 3xFP16 Gaussian image, when the model's range allows it (else 0) */
    void* b_h16 = h16_bytes ? slurp(f, h16_bytes) : NULL;
    int64_t* off = slurp(f, (size_t)(n_utt + 1) * 8);
    const size_t pcm_bytes = (size_t)n_samples * (pcm_format == LOE_PCM_S16 ? 2 : 4);
    void* pcm = NULL;
    if (loe_host_alloc(&pcm, pcm_bytes) != LOE_OK) { fprintf(stderr, "%s\n", loe_last_error()); return 1; }
    if (fread(pcm, 1, pcm_bytes, f) != pcm_bytes) { fprintf(stderr, "short pcm\n"); return 2; }
    fclose(f);

    void* dec = NULL;
    if (loe_decoder_create(0, mel_bin, mel_w, mel_na, mel_nb, b_packed, cst_pad, n_states, n_pos, col, band, flags, word, word_lo,
                           &dec) != LOE_OK) { fprintf(stderr, "create: %s\n", loe_last_error()); return 1; }
    if (b_h16 && h16_bytes == (size_t)n_tiles * (size_t)loe_emission_h16_tile_bytes() &&
        loe_decoder_set_h16(dec, b_h16) != LOE_OK) { fprintf(stderr, "set_h16: %s\n", loe_last_error()); return 1; }
    int8_t* words = malloc((size_t)n_utt * max_words);
    int32_t* count = malloc((size_t)n_utt * 4);
    float* score = malloc((size_t)n_utt * 4);
    for (int rep = 0; rep < 2; ++rep)           /* the second call reuses the workspace */
        if (loe_decoder_decode_host(dec, pcm, pcm_format, off, n_utt, penalty, penalty_f64, skip_label, max_words, n_chunks,
                                    words, count, score, NULL) != LOE_OK) { fprintf(stderr, "decode: %s\n", loe_last_error()); return 1; }
    for (int i = 0; i < n_utt; ++i) {
        printf("%d %.9g", count[i], score[i]);
        for (int k = 0; k < count[i] && k < max_words; ++k) printf(" %d", words[(size_t)i * max_words + k]);
        printf("\n");
    }
    loe_decoder_destroy(dec);
    loe_host_free(pcm);
    return 0;
}
