"""Base-state -> input-plane description (host side).

Builds, from a game's base propositions and its ``GameDesc``, the two tables the encoder needs:
``board_space`` (base -> channel, x, y) and ``control_space`` (base -> control channel, value), and
offers a numpy reference encoder. Same construction rules and attribute names as the reference's
Python ``GdlBasesTransformer`` (``src/ggpzero/nn/bases.py:62-287``): board channel ids are assigned
in first-seen order while walking the base list, ``x`` indexes ``x_cords`` (the reference calls its
count ``num_rows``) and ``y`` indexes ``y_cords``; planes are channels-first
``[channel][y_idx][x_idx]`` with ``buf offset = channel_size*channel + y_idx*num_rows + x_idx``
(``src/ggpzero/util/cppinterface.py:44``).
"""
import numpy as np

from .defs import datadesc, gamedesc
from .gdl.parser import parse_sexprs


class BaseInfo(object):
    def __init__(self, index, base_str):
        self.index = index
        sx = parse_sexprs(base_str)[0]           # ('true', <term>)
        t = sx[1]
        self.terms = list(t) if isinstance(t, tuple) else [t]
        self.used = False


class BaseToBoardSpace(object):
    def __init__(self, base_indx, channel_id, x_idx, y_idx):
        self.base_indx, self.channel_id, self.x_idx, self.y_idx = base_indx, channel_id, x_idx, y_idx

    def __repr__(self):
        return "BaseToBoardSpace(bs=%s, c=%s, x=%s, y=%s)" % (self.base_indx, self.channel_id,
                                                              self.x_idx, self.y_idx)


class BaseToChannelSpace(object):
    def __init__(self, base_indx, channel_id, value):
        self.base_indx, self.channel_id, self.value = base_indx, channel_id, value

    def __repr__(self):
        return "BaseToChannelSpace(b=%s, c=%s, v=%s)" % (self.base_indx, self.channel_id, self.value)


class GdlBasesTransformer(object):
    def __init__(self, game_info, generation_descr, game_desc=None, verbose=False):
        assert isinstance(generation_descr, datadesc.GenerationDescription)
        self.game_info = game_info
        if game_desc is None:
            game_desc = getattr(gamedesc.Games(), self.game)()
        self.game_desc = game_desc
        model = game_info.model
        self.role_count = len(model.roles)
        assert generation_descr.multiple_policy_heads
        self.policy_dist_count = [len(a) for a in model.actions]
        self.final_score_count = len(model.roles)
        self.channel_last = generation_descr.channel_last
        self.num_previous_states = generation_descr.num_previous_states
        self.num_rewards = 3 if generation_descr.draw_head else 2
        assert self.num_previous_states >= 0
        self.init_spaces()

    # ---- shape properties (names as in the reference, including its rows/cols naming)
    @property
    def game(self):
        return self.game_info.game

    @property
    def x_cords(self):
        return self.game_desc.x_cords

    @property
    def y_cords(self):
        return self.game_desc.y_cords

    @property
    def num_rows(self):
        return len(self.x_cords)

    @property
    def num_cols(self):
        return len(self.y_cords)

    @property
    def channel_size(self):
        return self.num_cols * self.num_rows

    @property
    def num_bases(self):
        return len(self.game_info.model.bases)

    @property
    def num_channels(self):
        return self.num_of_controls_channels + self.raw_channels_per_state * (self.num_previous_states + 1)

    # ---- table construction
    def init_spaces(self):
        infos = [BaseInfo(i, s) for i, s in enumerate(self.game_info.model.bases)]
        self.board_space = self.create_board_space(infos)
        self.raw_channels_per_state = max(b.channel_id for b in self.board_space) + 1
        self.control_space = self.create_control_space(infos)
        self.num_of_controls_channels = len(self.game_desc.control_channels)
        self.num_unhandled_states = sum(1 for b in infos if not b.used)

    def create_board_space(self, infos):
        space = []
        channel_of = {}
        for info in infos:
            bc = None
            for cand in self.game_desc.board_channels:
                if info.terms[0] == cand.base_term:
                    bc = cand
                    break
            if bc is None:
                continue
            matched = []
            for bt in bc.board_terms:
                if bt.term_idx < len(info.terms) and info.terms[bt.term_idx] in bt.terms:
                    matched.append(info.terms[bt.term_idx])
            if len(matched) != len(bc.board_terms):
                continue
            key = tuple([info.terms[0]] + matched)
            if key not in channel_of:
                channel_of[key] = len(channel_of)
            x_idx = self.game_desc.x_cords.index(info.terms[bc.x_term_idx])
            y_idx = self.game_desc.y_cords.index(info.terms[bc.y_term_idx])
            space.append(BaseToBoardSpace(info.index, channel_of[key], x_idx, y_idx))
            info.used = True
        return space

    def create_control_space(self, infos):
        space = []
        for channel_id, cc in enumerate(self.game_desc.control_channels):
            wanted = [tuple(str(t) for t in cb.arg_terms) for cb in cc.control_bases]
            for info in infos:
                terms = tuple(info.terms)
                if terms in wanted:
                    cb = cc.control_bases[wanted.index(terms)]
                    assert not info.used
                    space.append(BaseToChannelSpace(info.index, channel_id, cb.value))
                    info.used = True
        return space

    # ---- numpy reference encoder (restates nn/bases.py:242-287; used by tests / tooling only)
    def state_to_channels(self, state, prev_states=None):
        prev_states = prev_states or []
        assert len(prev_states) <= self.num_previous_states
        ch = np.zeros((self.num_channels, self.num_cols, self.num_rows), dtype=np.float32)
        for b in self.board_space:
            if state[b.base_indx]:
                ch[b.channel_id][b.y_idx][b.x_idx] = 1
        incr = self.raw_channels_per_state
        for ii in range(self.num_previous_states):
            if ii < len(prev_states):
                ps = prev_states[ii]
                for b in self.board_space:
                    if ps[b.base_indx]:
                        ch[b.channel_id + incr][b.y_idx][b.x_idx] = 1
            incr += self.raw_channels_per_state
        for c in self.control_space:
            if state[c.base_indx]:
                ch[c.channel_id + incr] += c.value
        if self.channel_last:
            ch = np.transpose(ch, (1, 2, 0))
        return ch

    # ---- sample -> network output rows. Same names and results as nn/bases.py:289-321 (the draw head's value row is
    # GdlBasesTransformer_Draws.value_to_array, :323-335 — chosen here by generation_descr.draw_head); written over
    # numpy arrays. tests/test_reference_python_golden.py holds them to the reference's own outputs.
    def check_sample(self, sample):
        """Raises AssertionError for a sample the reference's check rejects; returns the sample otherwise."""
        from .gendata import decode_state
        assert isinstance(sample, datadesc.Sample)
        state = decode_state(sample.state) if isinstance(sample.state, str) else sample.state
        assert len(state) >= self.num_bases, 'state shorter than the base count'      # (>=: the codec pads to bytes)
        assert len(sample.final_score) == self.final_score_count
        for policy in sample.policies:
            probs = np.array([p for _, p in policy], dtype=np.float64)
            assert np.all((probs > -0.01) & (probs < 1.01)), 'probability out of range'
            total = 0.0
            for p in probs:                      # (left to right, as the reference adds them)
                total += float(p)
            assert 0.99 < total < 1.01, 'probabilities do not sum to one'
        return sample

    def policy_to_array(self, policy, role_index):
        array = np.zeros(self.policy_dist_count[role_index], dtype=np.float32)
        if len(policy):
            idx, probs = zip(*policy)
            array[np.asarray(idx, dtype=np.int64)] = np.asarray(probs, dtype=np.float32)   # (a repeated index: the last wins)
        return array

    def value_to_array(self, values):
        values = [float(v) for v in values]
        if self.num_rewards != 3:
            assert len(values) == self.role_count
            return np.asarray(values, dtype=np.float32)
        # draw head: (win 0, win 1, draw) — a 0.5 / 0.5 result is the draw class
        assert len(values) == 2 and abs(values[0] + values[1] - 1.0) < 0.01
        drawn = abs(values[0] - 0.5) < 0.01
        assert not drawn or abs(values[1] - 0.5) < 0.01
        return np.asarray([0.0, 0.0, 1.0] if drawn else [values[0], values[1], 0.0], dtype=np.float32)

    def get_symmetries_desc(self):
        return getattr(gamedesc.GameSymmetries(), self.game, lambda: None)()

    def encode_tables(self):
        """(board_base, board_offset, control_base, control_channel, control_value) arrays in the
        order ``create_c_transformer`` feeds them to the engine (cppinterface.py:33-50)."""
        bb = np.array([b.base_indx for b in self.board_space], dtype=np.int32)
        bo = np.array([self.channel_size * b.channel_id + b.y_idx * self.num_rows + b.x_idx
                       for b in self.board_space], dtype=np.int32)
        cb = np.array([c.base_indx for c in self.control_space], dtype=np.int32)
        cc = np.array([c.channel_id for c in self.control_space], dtype=np.int32)
        cv = np.array([c.value for c in self.control_space], dtype=np.float32)
        return bb, bo, cb, cc, cv


GdlBasesTransformer_Draws = GdlBasesTransformer      # (the value row follows generation_descr.draw_head)
