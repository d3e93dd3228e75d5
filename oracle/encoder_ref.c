/* ORACLE — test infrastructure (see oracle/propnet_ref.c header for who may use it).
 *
 * Plain-C restatement of the reference's base-state -> input-plane encoder:
 *   GdlBasesTransformer::setForState   src/cpp/gdltransformer.cpp:14-20
 *   GdlBasesTransformer::toChannels    src/cpp/gdltransformer.cpp:22-52
 *   BaseToBoardSpace / BaseToChannelSpace   src/cpp/gdltransformer.h:15-66
 *   totalSize / controlStatesStart     src/cpp/gdltransformer.h:97-109
 * including its quirks: `count` is never incremented, so every previous state lands in group 1
 * (gdltransformer.cpp:39-43); control channels are flood-filled (overwritten, last set base wins).
 * This C++ file of the reference compiles on its own only against ggplib's BaseState header, which
 * is absent; BaseState::get(i) is restated as the LSB-first bit test used throughout this repo.
 * PARITY STATUS: pinned by code: the algorithm is fully specified by the lines
 * above; no golden plane dumps exist in the reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int channel_size, channels_per_state, num_control_channels, num_prev_states;
    int n_board, n_control, cap_board, cap_control;
    int *board_base, *board_incr;
    int *control_base, *control_channel;
    float* control_value;
} encref;

encref* encref_create(int channel_size, int channels_per_state, int num_control_channels,
                      int num_prev_states) {
    encref* t = (encref*)calloc(1, sizeof(encref));
    t->channel_size = channel_size;
    t->channels_per_state = channels_per_state;
    t->num_control_channels = num_control_channels;
    t->num_prev_states = num_prev_states;
    return t;
}

void encref_destroy(encref* t) {
    if (!t) return;
    free(t->board_base); free(t->board_incr); free(t->control_base); free(t->control_channel);
    free(t->control_value); free(t);
}

void encref_add_board_base(encref* t, int base_indx, int buf_incr) {
    if (t->n_board == t->cap_board) {
        t->cap_board = t->cap_board ? 2 * t->cap_board : 256;
        t->board_base = (int*)realloc(t->board_base, sizeof(int) * t->cap_board);
        t->board_incr = (int*)realloc(t->board_incr, sizeof(int) * t->cap_board);
    }
    t->board_base[t->n_board] = base_indx;
    t->board_incr[t->n_board++] = buf_incr;
}

void encref_add_control_base(encref* t, int base_indx, int channel_id, float value) {
    if (t->n_control == t->cap_control) {
        t->cap_control = t->cap_control ? 2 * t->cap_control : 64;
        t->control_base = (int*)realloc(t->control_base, sizeof(int) * t->cap_control);
        t->control_channel = (int*)realloc(t->control_channel, sizeof(int) * t->cap_control);
        t->control_value = (float*)realloc(t->control_value, sizeof(float) * t->cap_control);
    }
    t->control_base[t->n_control] = base_indx;
    t->control_channel[t->n_control] = channel_id;
    t->control_value[t->n_control++] = value;
}

int encref_total_size(const encref* t) {
    return t->channel_size * (t->channels_per_state * (t->num_prev_states + 1) + t->num_control_channels);
}

static int get_bit(const uint32_t* s, int b) { return (s[b >> 5] >> (b & 31)) & 1; }

static void set_for_state(const encref* t, float* local_buf, const uint32_t* bs) {
    for (int i = 0; i < t->n_board; i++)
        if (get_bit(bs, t->board_base[i])) local_buf[t->board_incr[i]] = 1.0f;
}

/* prev_states: n_prev pointers (nearest ancestor first); buf: total_size floats */
void encref_to_channels(const encref* t, const uint32_t* state, const uint32_t* const* prev_states,
                        int n_prev, float* buf) {
    const int total = encref_total_size(t);
    for (int i = 0; i < total; i++) buf[i] = 0.0f;
    set_for_state(t, buf, state);
    const int count = 1; /* never incremented in the reference */
    for (int p = 0; p < n_prev; p++) {
        float* local_buf = buf + t->channels_per_state * t->channel_size * count;
        set_for_state(t, local_buf, prev_states[p]);
    }
    float* control_start = buf + t->channel_size * (t->channels_per_state * (t->num_prev_states + 1));
    for (int i = 0; i < t->n_control; i++) {
        if (get_bit(state, t->control_base[i])) {
            float* b = control_start + t->channel_size * t->control_channel[i];
            for (int k = 0; k < t->channel_size; k++) b[k] = t->control_value[i];
        }
    }
}

/* batch helper: states [n][sw]; prev [n][max_prev][sw] with prev_count[n] valid entries */
void encref_batch(const encref* t, int64_t n, int sw, const uint32_t* states, const uint32_t* prev,
                  const uint8_t* prev_count, int max_prev, float* out) {
    const int total = encref_total_size(t);
    const uint32_t* ptrs[16];
    for (int64_t i = 0; i < n; i++) {
        int np = prev_count ? prev_count[i] : 0;
        if (np > 16) np = 16;
        for (int p = 0; p < np; p++) ptrs[p] = prev + ((size_t)i * max_prev + p) * sw;
        encref_to_channels(t, states + (size_t)i * sw, ptrs, np, out + (size_t)i * total);
    }
}
