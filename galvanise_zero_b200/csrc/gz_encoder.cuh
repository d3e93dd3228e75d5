// Device-side description of a GdlBasesTransformer and the per-element plane function shared by
// the stand-alone encoder kernel (gz_encoder.cu) and the PUCT leaf encoder (gz_puct.cu).
#ifndef GZ_ENCODER_CUH
#define GZ_ENCODER_CUH

#include <cuda_bf16.h>
#include <stdint.h>

struct EncParams {
    const int* src_base;       // [group_size] base index feeding plane cell (channel,y,x), -1 = none
    const int* ctrl_base;      // [n_control]
    const int* ctrl_channel;   // [n_control]
    const float* ctrl_value;   // [n_control]
    int n_control;
    int channel_size;
    int group_size;            // channels_per_state * channel_size
    int num_prev_states;
    int num_control_channels;
    int board_total;           // group_size * (num_prev_states + 1)
    int total;                 // board_total + num_control_channels * channel_size
};

__device__ __forceinline__ uint32_t enc_bit(const uint32_t* words, int b) {
    return (words[b >> 5] >> (b & 31)) & 1u;
}

// Value of control channel c for a state: the value of the LAST set control base in insertion order
// (flood fill overwrites, src/cpp/gdltransformer.cpp:46-51), 0 when none is set.
__device__ __forceinline__ float enc_control_value(const EncParams& E, const uint32_t* cur, int c) {
    float v = 0.0f;
    for (int i = 0; i < E.n_control; ++i)
        if (E.ctrl_channel[i] == c && enc_bit(cur, E.ctrl_base[i])) v = E.ctrl_value[i];
    return v;
}

// Plane element `le` (0 <= le < E.total) of one state. `prev_or` = OR of the available previous
// states (all of them land in group 1, gdltransformer.cpp:39-43), or nullptr.
__device__ __forceinline__ float enc_element(const EncParams& E, const uint32_t* cur,
                                             const uint32_t* prev_or, const float* ctrl, int le) {
    if (le < E.board_total) {
        const int g = le / E.group_size;
        const int l2 = le - g * E.group_size;
        const int b = E.src_base[l2];
        if (b < 0) return 0.0f;
        if (g == 0) return enc_bit(cur, b) ? 1.0f : 0.0f;
        if (g == 1 && prev_or) return enc_bit(prev_or, b) ? 1.0f : 0.0f;
        return 0.0f;
    }
    return ctrl[(le - E.board_total) / E.channel_size];
}

#endif
