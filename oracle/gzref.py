"""ORACLE (test infrastructure): ctypes binding of oracle/_ref/libgzref.so — the reference's own
PUCT / scheduler / self-play C++ compiled unmodified against shims (oracle/Makefile `ref`).
Only tests/, smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_ref', 'libgzref.so')
REF_SRC = '/root/reference/src/cpp'

PUCT_FIELDS = ['puct_constant', 'puct_constant_root', 'dirichlet_noise_pct', 'noise_policy_squash_pct',
               'noise_policy_squash_prob', 'choose', 'random_scale', 'temperature',
               'depth_temperature_start', 'depth_temperature_increment', 'depth_temperature_stop',
               'depth_temperature_max', 'fpu_prior_discount', 'fpu_prior_discount_root',
               'top_visits_best_guess_converge_ratio', 'think_time', 'converged_visits', 'batch_size',
               'use_legals_count_draw', 'backup_finalised', 'lookup_transpositions',
               'evaluation_multiplier_to_convergence']
SP_FIELDS = ['oscillate_sampling_pct', 'temperature_for_policy', 'evals_per_move',
             'resign0_score_probability', 'resign0_pct', 'resign1_score_probability', 'resign1_pct',
             'run_to_end_pct', 'run_to_end_evals', 'run_to_end_early_score',
             'run_to_end_minimum_game_depth', 'abort_max_length']


def available():
    return os.path.exists(LIB) or os.path.isdir(REF_SRC)


def build():
    """Builds oracle/_ref/libgzref.so when the reference sources are present (build container)."""
    if os.path.isdir(REF_SRC):
        subprocess.check_call(['make', '-s', '-C', HERE, 'ref'])
    return os.path.exists(LIB)


_L = None


def lib():
    global _L
    if _L is None:
        if not os.path.exists(LIB):
            if not build():
                raise RuntimeError('oracle/_ref/libgzref.so is missing and /root/reference is absent')
        L = ctypes.CDLL(LIB)
        vp, ci, dp = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
        L.gzref_create.restype = vp
        L.gzref_create.argtypes = [vp, ctypes.c_int64, ci, ci, ci, ci, ci, vp, ci]
        L.gzref_destroy.argtypes = [vp]
        L.gzref_add_board_base.argtypes = [vp, ci, ci]
        L.gzref_add_control_base.argtypes = [vp, ci, ci, ctypes.c_float]
        L.gzref_seed.argtypes = [ctypes.c_uint64]
        L.gzref_set_stagger.argtypes = [ci, ctypes.c_uint64]
        L.gzref_search.argtypes = [vp, dp, vp, ci, ci, ctypes.c_uint64, ci, vp, vp, vp, vp, vp, vp, vp]
        L.gzref_selfplay.argtypes = [vp, dp, dp, dp, ci, ci, ctypes.c_uint64, ctypes.c_int64,
                                     ctypes.c_double, dp, vp, ctypes.c_int64, ci]
        _L = L
    return _L


def puct_conf_array(conf):
    d = conf if isinstance(conf, dict) else __import__('attr').asdict(conf)
    vals = []
    for f in PUCT_FIELDS:
        v = d[f]
        if f == 'choose':
            v = 1.0 if v == 'choose_temperature' else 0.0
        vals.append(float(v))
    return (ctypes.c_double * len(vals))(*vals)


def sp_conf_array(conf):
    d = conf if isinstance(conf, dict) else __import__('attr').asdict(conf)
    return (ctypes.c_double * len(SP_FIELDS))(*[float(d[f]) for f in SP_FIELDS])


class RefEngine:
    def __init__(self, blob, c_transformer):
        """blob: tables.blob_by_name(game); c_transformer: interface.GdlBasesTransformer (tables)."""
        self.L = lib()
        self.blob = np.ascontiguousarray(blob, dtype=np.int32)
        ct = c_transformer
        ps = (ctypes.c_int * len(ct.policy_sizes))(*ct.policy_sizes)
        self.h = self.L.gzref_create(self.blob.ctypes.data, len(self.blob), ct.channel_size,
                                     ct.channels_per_state, ct.num_control_channels,
                                     ct.num_prev_states, ct.num_rewards, ps, len(ct.policy_sizes))
        assert self.h
        self.ct_total = ct.total_size()
        for b, off in ct.board_bases:
            self.L.gzref_add_board_base(self.h, b, off)
        for b, c, v in ct.control_bases:
            self.L.gzref_add_control_base(self.h, b, c, v)

    def encode(self, states, prev_states=None, prev_count=None):
        """The reference's compiled GdlBasesTransformer::toChannels on packed states uint32 [n, SW]
        (prev_states uint32 [n, max_prev, SW], nearest first; prev_count uint8 [n]) -> float32 [n, total]."""
        states = np.ascontiguousarray(states, dtype=np.uint32)
        n = states.shape[0]
        total = self.ct_total
        out = np.zeros((n, total), dtype=np.float32)
        pp = pc = None
        max_prev = 0
        if prev_states is not None:
            prev_states = np.ascontiguousarray(prev_states, dtype=np.uint32)
            max_prev = prev_states.shape[1]
            pp = prev_states.ctypes.data
            if prev_count is not None:
                prev_count = np.ascontiguousarray(prev_count, dtype=np.uint8)
                pc = prev_count.ctypes.data
        self.L.gzref_encode.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        got = self.L.gzref_encode(self.h, n, states.ctypes.data, pp, pc, max_prev, out.ctypes.data)
        assert got == total, (got, total)
        return out

    def search(self, puct_conf, evals, n_moves=1, root_state=None, nn_seed=1, max_children=512,
               rng_seed=0x5EED0000):
        self.L.gzref_seed(rng_seed)
        counts = np.zeros(n_moves, dtype=np.int32)
        move = np.full((n_moves, max_children), -1, dtype=np.int32)
        trav = np.zeros((n_moves, max_children), dtype=np.int32)
        prob = np.zeros((n_moves, max_children), dtype=np.float32)
        score = np.zeros((n_moves, max_children), dtype=np.float32)
        choice = np.zeros(n_moves, dtype=np.int32)
        root = np.zeros((n_moves, 4), dtype=np.float32)
        rs = None
        if root_state is not None:
            rs = np.ascontiguousarray(root_state, dtype=np.uint32)
        n = self.L.gzref_search(self.h, puct_conf_array(puct_conf), rs.ctypes.data if rs is not None else None,
                                evals, n_moves, nn_seed, max_children, counts.ctypes.data,
                                move.ctypes.data, trav.ctypes.data, prob.ctypes.data,
                                score.ctypes.data, choice.ctypes.data, root.ctypes.data)
        cs = np.full((max(n, 1), 4), -1.0, dtype=np.float32)
        self.L.gzref_last_choice_scores.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.L.gzref_last_choice_scores(cs.ctypes.data, cs.size)
        bal = np.zeros(max(n, 1) * (max_children + 1), dtype=np.int32)
        self.L.gzref_last_balance.argtypes = [ctypes.c_void_p, ctypes.c_int]
        nb = self.L.gzref_last_balance(bal.ctypes.data, bal.size)
        balance, pos = [], 0
        while pos < nb:
            k = int(bal[pos])
            balance.append(bal[pos + 1:pos + 1 + k].copy())
            pos += 1 + k
        dbg = np.zeros((max(n, 1), 3, 25), dtype=np.float32)
        self.L.gzref_last_debug.argtypes = [ctypes.c_void_p, ctypes.c_int]
        self.L.gzref_last_debug(dbg.ctypes.data, dbg.size)
        out = []
        for m in range(n):
            k = int(counts[m])
            lines = []
            for ln in dbg[m]:
                if ln[0] > 0:       # (an unexpanded child leaves its line empty, node.cpp:412-414)
                    nv = int(ln[4])
                    lines.append((float(ln[1]), int(ln[2]), int(ln[3]),
                                  tuple((int(ln[5 + 2 * v]), int(ln[6 + 2 * v])) for v in range(nv))))
            out.append({'moves': move[m, :k].copy(), 'traversals': trav[m, :k].copy(),
                        'choice_scores': cs[m].copy(), 'tree_debug': tuple(lines),
                        'balanced_traversals': balance[m] if m < len(balance) else None,
                        'policy_prob': prob[m, :k].copy(), 'scores': score[m, :k].copy(),
                        'choice': int(choice[m]), 'root_visits': int(root[m, 0]),
                        'lead_role': int(root[m, 1]), 'root_score': float(root[m, 2]),
                        'nodes': int(root[m, 3])})
        return out

    def set_balance(self, max_moves):
        """Every search of ``search`` is followed by ``balanceFirstMoves(max_moves)`` (0: off)."""
        self.L.gzref_set_balance.argtypes = [ctypes.c_int]
        self.L.gzref_set_balance(int(max_moves))

    def selfplay(self, sp_conf, batch_size, num_workers, nn_seed=1, target_samples=0,
                 max_seconds=10.0, rng_seed=0x5EED0000, sample_capacity=0, srec=0, warm_seconds=0.0,
                 stagger_max_ply=0, stagger_seed=0xB200):
        """Runs the reference Supervisor. With sample_capacity > 0 also returns the samples as an
        int32 array [n, srec] in the record layout of csrc/gz_puct_core.h (SR_*). stagger_max_ply > 0:
        the first game of every slot starts a uniformly random number of random plies (< max) into a game."""
        self.L.gzref_seed(rng_seed)
        self.L.gzref_set_stagger(int(stagger_max_ply), int(stagger_seed))
        import attr
        d = sp_conf if isinstance(sp_conf, dict) else attr.asdict(sp_conf)
        stats = (ctypes.c_double * 12)()
        stats[8] = float(warm_seconds)
        buf = np.zeros((max(sample_capacity, 1), max(srec, 1)), dtype=np.int32)
        self.L.gzref_selfplay(self.h, sp_conf_array(d), puct_conf_array(d['puct_config']),
                              puct_conf_array(d['run_to_end_puct_config']), batch_size, num_workers,
                              nn_seed, target_samples, max_seconds, stats,
                              buf.ctypes.data if sample_capacity else None, sample_capacity, srec)
        keys = ['samples', 'evaluations', 'polls', 'seconds', 'depth_sum', 'sampled_moves', 'games',
                'plies', 'moves_total', 'window_seconds', 'window_moves', 'window_evals']
        res = dict(zip(keys, [float(x) for x in stats]))
        if sample_capacity:
            res['records'] = buf[:min(int(res['samples']), sample_capacity)]
        return res
