"""ctypes driver for tests/hostsim/libhostsim.so (CPU validation build of the PUCT core)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'hostsim')
_L = None


def lib():
    global _L
    if _L is None:
        subprocess.check_call(['make', '-s', '-C', HERE])
        L = ctypes.CDLL(os.path.join(HERE, 'libhostsim.so'))
        vp, ci, dp = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double)
        L.hs_create.restype = vp
        L.hs_create.argtypes = [vp, ctypes.c_int64, ci, ci, ci, ci, ci, ci, ci, ci, ci, vp, ci, vp, vp, ci,
                                dp, dp, dp, ctypes.c_uint64, ctypes.c_uint64]
        L.hs_destroy.argtypes = [vp]
        L.hs_step.argtypes = [vp]
        L.hs_cmd_reset.argtypes = [vp, ci, vp, ci]
        L.hs_cmd_search.argtypes = [vp, ci, ci]
        L.hs_cmd_apply.argtypes = [vp, ci, ci]
        L.hs_cmd_balance.argtypes = [vp, ci, ci]
        L.hs_wait.argtypes = [vp, ci, ci]
        L.hs_root_children.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
        L.hs_ready_count.restype = ctypes.c_ulonglong
        L.hs_ready_count.argtypes = [vp]
        L.hs_ready.restype = ctypes.POINTER(ctypes.c_int32)
        L.hs_ready.argtypes = [vp]
        L.hs_srec.argtypes = [vp]
        L.hs_set_answer_cap.argtypes = [vp, ci]
        L.hs_transpositions.restype = ctypes.c_ulonglong
        L.hs_transpositions.argtypes = [vp, ci]
        L.hs_ctl.restype = vp
        L.hs_ctl.argtypes = [vp, ci]
        _L = L
    return _L


class HostSim:
    def __init__(self, blob, ct, sp_conf, G=1, mode=1, ncap=8192, ccap=1 << 17, nn_seed=1,
                 rng_seed=0x5EED0000):
        from oracle.gzref import puct_conf_array, sp_conf_array
        import attr
        self.L = lib()
        d = sp_conf if isinstance(sp_conf, dict) else attr.asdict(sp_conf)
        self.blob = np.ascontiguousarray(blob, dtype=np.int32)
        board = np.array(ct.board_bases, dtype=np.int32).reshape(-1, 2)
        cbc = np.array([(b, c) for b, c, v in ct.control_bases], dtype=np.int32).reshape(-1, 2)
        cv = np.array([v for b, c, v in ct.control_bases], dtype=np.float32)
        self._keep = (board, cbc, cv)
        self.h = self.L.hs_create(self.blob.ctypes.data, len(self.blob), G, mode, ncap, ccap,
                                  ct.channel_size, ct.channels_per_state, ct.num_control_channels,
                                  ct.num_prev_states, ct.num_rewards, board.ctypes.data, len(board),
                                  cbc.ctypes.data, cv.ctypes.data, len(cv), sp_conf_array(d),
                                  puct_conf_array(d['puct_config']),
                                  puct_conf_array(d['run_to_end_puct_config']), nn_seed, rng_seed)
        assert self.h

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.hs_destroy(self.h)
            self.h = None

    def search(self, evals, n_moves, root_state=None, g=0, max_children=512, balance=0):
        L = self.L
        rs = np.ascontiguousarray(root_state, dtype=np.uint32) if root_state is not None else None
        L.hs_cmd_reset(self.h, g, rs.ctypes.data if rs is not None else None, 0)
        assert L.hs_wait(self.h, g, 100000) == 1
        out = []
        for m in range(n_moves):
            probe = np.zeros(8, dtype=np.float32)
            tmp_i = np.zeros(max_children, dtype=np.int32)
            tmp_f = np.zeros(max_children, dtype=np.float32)
            if L.hs_root_children(self.h, g, max_children, tmp_i.ctypes.data, tmp_i.ctypes.data,
                                  tmp_f.ctypes.data, tmp_f.ctypes.data, probe.ctypes.data) == 0:
                break       # terminal root
            L.hs_cmd_search(self.h, g, evals)
            rc = L.hs_wait(self.h, g, 10000000)
            assert rc == 1, rc
            moves = np.zeros(max_children, dtype=np.int32)
            trav = np.zeros(max_children, dtype=np.int32)
            prob = np.zeros(max_children, dtype=np.float32)
            score = np.zeros(max_children, dtype=np.float32)
            info = np.zeros(8, dtype=np.float32)
            k = L.hs_root_children(self.h, g, max_children, moves.ctypes.data, trav.ctypes.data,
                                   prob.ctypes.data, score.ctypes.data, info.ctypes.data)
            choice = int(info[4])
            balanced = None
            if balance > 0:
                L.hs_cmd_balance(self.h, g, balance)
                assert L.hs_wait(self.h, g, 10000000) == 1
                bt = np.zeros(max_children, dtype=np.int32)
                L.hs_root_children(self.h, g, max_children, tmp_i.ctypes.data, bt.ctypes.data, tmp_f.ctypes.data,
                                   tmp_f.ctypes.data, probe.ctypes.data)
                balanced = bt[:k].copy()
            out.append({'moves': moves[:k].copy(), 'traversals': trav[:k].copy(), 'balanced_traversals': balanced,
                        'policy_prob': prob[:k].copy(), 'scores': score[:k].copy(), 'choice': choice,
                        'root_visits': int(info[0]), 'lead_role': int(info[1]),
                        'root_score': float(info[2]), 'nodes': int(info[3])})
            L.hs_cmd_apply(self.h, g, int(moves[choice]))
            rc = L.hs_wait(self.h, g, 100000)
            if rc != 1:
                break
            if os.environ.get('GZREF_TRACE'):
                import sys
                ti = np.zeros(max_children, dtype=np.int32); tf = np.zeros(max_children, dtype=np.float32); inf = np.zeros(8, dtype=np.float32)
                L.hs_root_children(self.h, g, max_children, ti.ctypes.data, ti.ctypes.data, tf.ctypes.data, tf.ctypes.data, inf.ctypes.data)
                print('A %d nodes %d root_visits %d' % (m, int(inf[3]), int(inf[0])), file=sys.stderr)
        return out

    def set_answer_cap(self, cap):
        self.L.hs_set_answer_cap(self.h, int(cap))

    def transpositions(self, g=0):
        return int(self.L.hs_transpositions(self.h, g))

    def step(self, n=1):
        for _ in range(n):
            self.L.hs_step(self.h)
