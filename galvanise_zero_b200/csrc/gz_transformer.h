// Host-side transformer object shared by gz_encoder.cu and gz_puct.cu.
#ifndef GZ_TRANSFORMER_H
#define GZ_TRANSFORMER_H

#include <vector>

#include "gz_encoder.cuh"

struct gz_transformer {
    int channel_size = 0, channels_per_state = 0, num_control_channels = 0, num_prev_states = 0;
    int num_rewards = 0;
    std::vector<int> policy_sizes;
    std::vector<int> board_base, board_off, ctrl_base, ctrl_channel;
    std::vector<float> ctrl_value;
    bool finalized = false;
    int device = 0, n_bases = 0, n_sms = 0;
    void* d_arena = nullptr;
    EncParams E{};
};

#endif
