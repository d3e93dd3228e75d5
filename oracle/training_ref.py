"""CPU restatement of the reference trainer's per-sample work — TEST INFRASTRUCTURE ONLY (the checker for
gz_records_to_training / galvanise_zero_b200.datacache; never imported by the product path).

Follows, sample by sample in plain Python / numpy:
  * Supervisor.fetch_samples' dict layout (src/cpp/pyobjects/supervisor_impl.cpp:75-118) for decoding a record,
  * DataCache.augment_data (src/ggpzero/nn/datacache.py:597-671): prescription loop, `seen` set, the uniform draw
    taken only when `seen` is non-empty, translate_basestate / translate_action per entry,
  * DataCache.sync (src/ggpzero/nn/datacache.py:714-760): final_score clamped with min(1.0, v) then max(0.0, v),
    state_to_channels, policy_to_array per role, value_to_array,
  * GdlBasesTransformer.policy_to_array / value_to_array / check_sample (src/ggpzero/nn/bases.py:289-314) and
    GdlBasesTransformer_Draws.value_to_array (:323-335).
Parity pin: the reference holds no stored training arrays; this restatement is pinned by the
code cited above and cross-checked in tests/test_training_batch.py against the Python transformer mirror
(galvanise_zero_b200/bases.py) fed with decoded sample dicts — two routes that share no code. Planes come from
the C oracle encoder (oracle/encoder_ref.c), itself pinned by the reference's compiled toChannels.
"""
import numpy as np

SR_DEPTH, SR_GAME_LENGTH, SR_SLOT, SR_MATCH, SR_HAS_RESIGNED, SR_RESIGN_FP, SR_START_DEPTH, SR_VISITS, \
    SR_LEAD, SR_N_PREV, SR_N_CHILDREN = range(1, 12)
SR_FINAL_SCORE, SR_PUCT_SCORE, SR_OTHER_MOVES, SR_HEADER = 12, 16, 20, 24


def unpack_bits(words, n_bits):
    w = np.ascontiguousarray(words, dtype=np.uint32)
    bits = np.unpackbits(w.view(np.uint8), bitorder='little')
    return bits[:n_bits].astype(np.uint8)


def pack_bits(bits, n_words):
    out = np.zeros(n_words * 32, dtype=np.uint8)
    out[:len(bits)] = np.asarray(bits, dtype=np.uint8)
    return np.packbits(out, bitorder='little').view(np.uint32)


def decode_record(rec, state_words, n_bases, n_roles):
    """One int32 record -> (state bits, [prev state bits], policies [[(legal, prob)]], final_score [float])."""
    rec = np.asarray(rec, dtype=np.int32)
    f = rec.view(np.float32)
    sw = state_words
    state = unpack_bits(rec[SR_HEADER:SR_HEADER + sw].view(np.uint32), n_bases)
    prevs = []
    if rec[SR_N_PREV]:
        prevs.append(unpack_bits(rec[SR_HEADER + sw:SR_HEADER + 2 * sw].view(np.uint32), n_bases))
    nc = int(rec[SR_N_CHILDREN])
    lead = int(rec[SR_LEAD])
    pairs = rec[SR_HEADER + 2 * sw:SR_HEADER + 2 * sw + 2 * nc].reshape(nc, 2)
    policies = []
    for ri in range(n_roles):
        if ri == lead:
            policies.append([(int(m), float(p)) for m, p in zip(pairs[:, 0], pairs[:, 1].copy().view(np.float32))])
        else:
            policies.append([(int(rec[SR_OTHER_MOVES + ri]), 1.0)])
    final_score = [float(f[SR_FINAL_SCORE + r]) for r in range(n_roles)]
    return state, prevs, policies, final_score


def translate_state(bits, base_dest):
    """Translator.translate_basestate (src/ggpzero/util/symmetry.py:178-198) with the per-base destination table."""
    out = np.zeros(len(bits), dtype=np.uint8)
    for i, b in enumerate(bits):
        if b:
            out[base_dest[i]] = 1
    return out


def check_sample(policies, final_score, widths, n_values):
    """bases.py:289-303 (+ the asserts of the draw head's value_to_array, :324-331, + index range): True if fine."""
    for ri, policy in enumerate(policies):
        total = 0.0
        for legal, p in policy:
            if not (0 <= legal < widths[ri]):
                return False
            if not (-0.01 < p < 1.01):
                return False
            total += p
        if not (0.99 < total < 1.01):
            return False
    if n_values == 3:
        if not abs(sum(final_score) - 1.0) < 0.01:
            return False
        if abs(final_score[0] - 0.5) < 0.01 and not abs(final_score[1] - 0.5) < 0.01:
            return False
    return True


def value_to_array(values, n_values):
    if n_values == 3:                                   # GdlBasesTransformer_Draws.value_to_array
        if abs(values[0] - 0.5) < 0.01:
            return np.array([0.0, 0.0, 1.0], dtype='float32')
        return np.array([values[0], values[1], 0.0], dtype='float32')
    return np.array(values, dtype='float32')


def policy_to_array(policy, width):
    array = np.zeros(width, dtype='float32')
    for idx, prob in policy:
        array[idx] = prob
    return array


def augmentation_rows(states, prescription_tables, draws, pct):
    """augment_data's control flow (datacache.py:629-653) -> [(sample, entry)] in the order it yields.
    states: list of bit arrays; prescription_tables: per entry the base destination table (entry 0 = identity);
    draws[i]: the uniform numbers sample i consumes, one per entry tried while `seen` is non-empty."""
    rows = []
    for i, bits in enumerate(states):
        seen = set()
        it = iter(draws[i])
        for k, dest in enumerate(prescription_tables):
            if len(seen) and next(it) > pct:
                continue
            state = tuple(translate_state(bits, dest).tolist())
            if state in seen:
                continue
            seen.add(state)
            rows.append((i, k))
    return rows


def records_to_training(records, state_words, n_bases, widths, n_values, encoder, rows=None, base_dest=None,
                        action_dest=None):
    """-> (planes [n, total] float32, [policy arrays float32 [n, A_r]], values float32 [n, n_values], bad count).
    rows: [(record, entry)] (None: every record once, untranslated); base_dest[k] / action_dest[r][k]: the
    destination tables of prescription entry k; encoder: oracle RefEncoder."""
    n_roles = len(widths)
    if rows is None:
        rows = [(i, None) for i in range(len(records))]
    n = len(rows)
    states = np.zeros((n, state_words), dtype=np.uint32)
    prevs = np.zeros((n, 1, state_words), dtype=np.uint32)
    cnt = np.zeros(n, dtype=np.uint8)
    pols = [np.zeros((n, w), dtype=np.float32) for w in widths]
    vals = np.zeros((n, n_values), dtype=np.float32)
    bad = 0
    for o, (i, k) in enumerate(rows):
        state, prev_states, policies, final_score = decode_record(records[i], state_words, n_bases, n_roles)
        if k is not None:
            state = translate_state(state, base_dest[k])
            prev_states = [translate_state(p, base_dest[k]) for p in prev_states]
        final_score = [min(1.0, v) for v in final_score]
        final_score = [max(0.0, v) for v in final_score]
        ok = check_sample(policies, final_score, widths, n_values)
        if k is not None:
            policies = [[(int(action_dest[ri][k][legal]), p) for legal, p in policy if 0 <= legal < widths[ri]]
                        for ri, policy in enumerate(policies)]
        else:
            policies = [[(legal, p) for legal, p in policy if 0 <= legal < widths[ri]]
                        for ri, policy in enumerate(policies)]
        if not ok:
            bad += 1
        states[o] = pack_bits(state, state_words)
        if prev_states:
            prevs[o, 0] = pack_bits(prev_states[0], state_words)
            cnt[o] = 1
        for ri in range(n_roles):
            pols[ri][o] = policy_to_array(policies[ri], widths[ri])
        vals[o] = value_to_array(final_score, n_values)
    planes = encoder.encode(states, prevs, cnt)
    return planes, pols, vals, bad
