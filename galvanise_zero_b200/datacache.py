"""Sample records -> training arrays on the device: the per-sample work of the reference trainer's
``DataCache.sync`` (``src/ggpzero/nn/datacache.py:700-760``: final scores clamped, ``state_to_channels``,
``policy_to_array`` per role, ``value_to_array``) and of ``DataCache.augment_data`` (``:597-671``: every sample
under the symmetries of the game's prescription, identical translated states dropped, the non-identity ones kept
with probability ``data_augment_pct``), for whole batches of the engine's sample records
(``PuctEngine.fetch_sample_records`` / ``distributed.gather_sample_records``) through
``gz_records_to_training`` (``csrc/gz_training.cu``). The reference's database (bcolz), chunk indexer and
statistics are the trainer's own and not rebuilt.

There is no CPU path here: without ``libgz_b200.so`` or a GPU the calls raise."""
import ctypes
import random

import numpy as np

from . import lib as _lib
from . import symmetry
from .defs import gamedesc


def _declare(L):
    if getattr(L, '_gz_trn_declared', False):
        return
    vp, ci, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.gz_records_to_training_scratch_bytes.argtypes = [i64, ci]
    L.gz_records_to_training_scratch_bytes.restype = i64
    L.gz_records_to_training.argtypes = [vp, i64, i64, ci, ci, ci, vp, vp, vp, ci, vp, vp, vp, ci, vp, vp, ci, vp,
                                         vp, vp]
    L.gz_permute_state_bits.argtypes = [i64, ci, ci, vp, vp, vp, vp]
    L._gz_trn_declared = True


class TrainingBatcher(object):
    """``arrays(records)`` -> planes, one dense policy array per role and the value array for a batch of records.

    ``transformer``: the Python ``bases.GdlBasesTransformer``; ``c_transformer``: its engine-side twin
    (``cppinterface.create_c_transformer``, finalised for ``device``); ``state_words``: words per packed state row.
    ``do_augment_data`` / ``data_augment_pct`` as in ``DataCache.__init__`` (``datacache.py:443-460``); the
    uniform draws come from ``rng`` (``random.Random``) in the reference's order: per sample, one draw per
    non-identity prescription entry."""

    def __init__(self, transformer, c_transformer, state_words, device=0, do_augment_data=False, data_augment_pct=1.0,
                 rng=None):
        import torch
        self._L = _lib.lib()
        _declare(self._L)
        self.transformer = transformer
        self.c_transformer = c_transformer.finalize(transformer.num_bases, device)
        self.device = torch.device('cuda', device)
        self.state_words = int(state_words)
        self.n_bases = int(transformer.num_bases)
        self.widths = [int(w) for w in transformer.policy_dist_count]
        self.n_values = int(transformer.num_rewards)
        self.plane_shape = (transformer.num_channels, transformer.num_cols, transformer.num_rows)
        self.do_augment_data = bool(do_augment_data)
        self.data_augment_pct = float(data_augment_pct)
        self.rng = rng or random.Random()
        self.prescription = [(False, 0)]
        self._base_src = self._action_dest = None
        desc = transformer.get_symmetries_desc()
        if self.do_augment_data and desc is not None:
            t = symmetry.create_translator(transformer.game_info, transformer.game_desc, desc)
            self.prescription = list(symmetry.Prescription(desc))
            assert self.prescription[0] == (False, 0)          # (datacache.py:635: the identity comes first)
            src = np.empty((len(self.prescription), self.n_bases), dtype=np.int32)
            for k, (refl, rot) in enumerate(self.prescription):
                dest = t.base_permutation(refl, rot)
                src[k, dest] = np.arange(self.n_bases, dtype=np.int32)     # new bit dest[i] = old bit i
            self._base_src = torch.from_numpy(src).to(self.device)
            self._action_dest = []
            for ri in range(len(self.widths)):
                d = np.stack([np.asarray(t.action_permutation(ri, refl, rot), dtype=np.int32)
                              for refl, rot in self.prescription])
                assert d.shape == (len(self.prescription), self.widths[ri])
                self._action_dest.append(torch.from_numpy(np.ascontiguousarray(d)).to(self.device))
            self._action_ptrs = (ctypes.c_void_p * len(self.widths))(*[a.data_ptr() for a in self._action_dest])

    # ---- augment_data's choice of (sample, symmetry) rows
    def _translated_states(self, records, k):
        import torch
        n, sw = records.shape[0], self.state_words
        states = records[:, _SR_HEADER:_SR_HEADER + sw].contiguous()
        out = torch.empty_like(states)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._L.gz_permute_state_bits(n, sw, self.n_bases, states.data_ptr(), self._base_src[k].data_ptr(),
                                                 out.data_ptr(), stream))
        return out

    def augmentation_rows(self, records):
        """(index, sym) int32 tensors: which record under which prescription entry makes each output row, in the
        order ``augment_data`` yields them (sample-major)."""
        import torch
        n = records.shape[0]
        K = len(self.prescription)
        if K == 1:
            return None, None
        keep = torch.ones((n, K), dtype=torch.bool, device=self.device)
        draws = np.array([[self.rng.random() for _ in range(K - 1)] for _ in range(n)], dtype=np.float64).reshape(n, K - 1)
        keep[:, 1:] = torch.from_numpy(draws <= self.data_augment_pct).to(self.device)    # `random() > pct`: skipped
        seen = [self._translated_states(records, 0)]
        for k in range(1, K):
            tk = self._translated_states(records, k)
            dup = torch.zeros(n, dtype=torch.bool, device=self.device)
            for j, prev in enumerate(seen):
                # a state counts as seen only if its own row was yielded (kept and not itself a duplicate)
                dup |= keep[:, j] & (prev == tk).all(dim=1)
            keep[:, k] &= ~dup
            seen.append(tk)
        idx = keep.nonzero()                                                               # row-major: sample-major
        return idx[:, 0].to(torch.int32).contiguous(), idx[:, 1].to(torch.int32).contiguous()

    def arrays_from_samples(self, samples, **kw):
        """``arrays`` for ``datadesc.Sample`` objects / fetch_samples dicts (e.g. ``gendata.read_gendata(path).samples``)."""
        import torch
        recs = samples_to_records(samples, self.transformer, self.state_words)
        return self.arrays(torch.from_numpy(recs).to(self.device), **kw)

    # ---- the arrays
    def arrays(self, records, plane_dtype=None, check=False, index=None, sym=None):
        """records: int32 CUDA tensor [n, record_words]. Returns a dict: ``planes`` [rows, C, H, W] (float32, or
        ``plane_dtype``), ``policies`` (list of float32 [rows, A_r]), ``value`` float32 [rows, num_rewards],
        ``index`` / ``sym`` (None without augmentation) and, with ``check``, ``bad`` = the number of rows that fail
        ``check_sample``."""
        import torch
        assert records.is_cuda and records.dtype == torch.int32 and records.is_contiguous() and records.dim() == 2
        plane_dtype = plane_dtype or torch.float32
        n_rec, rw = records.shape
        if index is None and self.do_augment_data and len(self.prescription) > 1:
            index, sym = self.augmentation_rows(records)
        n = n_rec if index is None else int(index.shape[0])
        if index is not None:
            assert index.dtype == torch.int32 and index.is_cuda and index.is_contiguous()
        if sym is not None:
            assert sym.dtype == torch.int32 and sym.is_cuda and sym.is_contiguous() and sym.shape[0] == n
            assert self._base_src is not None, 'symmetry ids need do_augment_data and a game with symmetries'
        planes = torch.empty((n,) + self.plane_shape, dtype=plane_dtype, device=self.device)
        policies = [torch.empty((n, w), dtype=torch.float32, device=self.device) for w in self.widths]
        value = torch.empty((n, self.n_values), dtype=torch.float32, device=self.device)
        bad = torch.zeros(1, dtype=torch.int32, device=self.device) if check else None
        scratch = torch.empty(max(1, int(self._L.gz_records_to_training_scratch_bytes(n, self.state_words))),
                              dtype=torch.uint8, device=self.device)
        pol_ptrs = (ctypes.c_void_p * len(policies))(*[p.data_ptr() for p in policies])
        code = {torch.float32: 0, torch.bfloat16: 1}[plane_dtype]
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self._L.gz_records_to_training(
            self.c_transformer._h, n, n_rec, rw, self.state_words, len(self.widths), records.data_ptr(),
            index.data_ptr() if index is not None else None, sym.data_ptr() if sym is not None else None,
            len(self.prescription), self._base_src.data_ptr() if sym is not None else None,
            self._action_ptrs if sym is not None else None, planes.data_ptr(), code, pol_ptrs, value.data_ptr(),
            self.n_values, bad.data_ptr() if check else None, scratch.data_ptr(), stream))
        out = dict(planes=planes, policies=policies, value=value, index=index, sym=sym)
        if check:
            out['bad'] = int(bad.item())
        return out


_SR_HEADER = 24        # include/gz_b200.h GZ_SR_HEADER
(_SR_DEPTH, _SR_GAME_LENGTH, _SR_SLOT, _SR_MATCH, _SR_HAS_RESIGNED, _SR_RESIGN_FP, _SR_START_DEPTH, _SR_VISITS, _SR_LEAD,
 _SR_N_PREV, _SR_N_CHILDREN) = range(1, 12)
_SR_FINAL_SCORE, _SR_PUCT_SCORE, _SR_OTHER_MOVES = 12, 16, 20


def samples_to_records(samples, transformer, state_words, record_words=None):
    """``datadesc.Sample`` objects (or the dicts of ``Supervisor.fetch_samples``) -> int32 sample records
    (include/gz_b200.h GZ_SR_*), the inverse of ``PuctEngine.records_to_samples`` — what lets a generation file
    written by the reference (``gendata_<game>_<step>.json.gz`` through ``gendata.read_gendata``) feed
    ``TrainingBatcher.arrays``. States may be bit lists or the reference's base64 strings (``util/state.py``). The
    role with the longest policy is the record's lead role (ties: the first); every other role must hold exactly
    one (move, probability) pair, as ``SelfPlayManager::createSample`` writes them (selfplaymanager.cpp:95-108)."""
    from . import tables
    from .gendata import decode_state
    widths = list(transformer.policy_dist_count)
    nb = transformer.num_bases
    rw = record_words or (_SR_HEADER + 2 * state_words + 2 * max(widths))
    recs = np.zeros((len(samples), rw), dtype=np.int32)
    f = recs.view(np.float32)
    get = (lambda s, k: s[k]) if samples and isinstance(samples[0], dict) else getattr
    for i, s in enumerate(samples):
        policies = get(s, 'policies')
        assert len(policies) == len(widths), 'one policy per role'
        lead = max(range(len(widths)), key=lambda ri: (len(policies[ri]), -ri))
        pairs = policies[lead]
        assert _SR_HEADER + 2 * state_words + 2 * len(pairs) <= rw, 'policy longer than a record holds'
        bits = np.asarray(decode_state(get(s, 'state')), dtype=np.uint8)[:nb]
        recs[i, _SR_HEADER:_SR_HEADER + state_words] = tables.pack_state_bits(bits, state_words).view(np.int32)
        prevs = get(s, 'prev_states')
        assert len(prevs) <= 1, 'sample records carry at most one previous state'
        if prevs:
            pb = np.asarray(decode_state(prevs[0]), dtype=np.uint8)[:nb]
            recs[i, _SR_HEADER + state_words:_SR_HEADER + 2 * state_words] = tables.pack_state_bits(pb, state_words).view(np.int32)
        recs[i, _SR_N_PREV] = len(prevs)
        recs[i, _SR_LEAD] = lead
        recs[i, _SR_N_CHILDREN] = len(pairs)
        p0 = _SR_HEADER + 2 * state_words
        recs[i, p0:p0 + 2 * len(pairs):2] = [int(m) for m, _ in pairs]
        f[i, p0 + 1:p0 + 2 * len(pairs):2] = [float(p) for _, p in pairs]
        for ri, pol in enumerate(policies):
            if ri != lead:
                assert len(pol) == 1, 'a role that is not on move has one (move, 1.0) pair'
            recs[i, _SR_OTHER_MOVES + ri] = int(pol[0][0])
        fs = get(s, 'final_score')
        ps = get(s, 'resultant_puct_score')
        for ri in range(len(widths)):
            f[i, _SR_FINAL_SCORE + ri] = fs[ri]
            if ri < len(ps):
                f[i, _SR_PUCT_SCORE + ri] = ps[ri]
        recs[i, _SR_DEPTH] = get(s, 'depth')
        recs[i, _SR_GAME_LENGTH] = get(s, 'game_length')
        recs[i, _SR_HAS_RESIGNED] = int(bool(get(s, 'has_resigned')))
        recs[i, _SR_RESIGN_FP] = int(bool(get(s, 'resign_false_positive')))
        recs[i, _SR_START_DEPTH] = get(s, 'starting_sample_depth')
        recs[i, _SR_VISITS] = get(s, 'resultant_puct_visits')
    return recs


def game_symmetries(game):
    """``GdlBasesTransformer.get_symmetries_desc`` by game name (``nn/bases.py:316-320``)."""
    return gamedesc.get_symmetries_desc(game)
