"""Python side of the engine boundary — mirrors ``src/ggpzero/util/cppinterface.py``:
``create_c_transformer`` (``:33-50``), ``PollerBase`` (``:53-153``), ``PlayPoller`` (``:156-182``)
and ``Supervisor`` (``:185-226``)."""
from __future__ import annotations

from . import interface as ggpzero_interface


def create_c_transformer(transformer, device=None, module=None):
    """Python ``GdlBasesTransformer`` (bases.py) -> engine-side transformer; same feeding order as
    the reference (``cppinterface.py:33-50``). With ``device`` given the tables are also uploaded (finalised).
    ``module``: the extension module to build it with (the compiled ``ggpzero_interface``; default: the ctypes
    host ``interface.py``)."""
    c_transformer = (module or ggpzero_interface).GdlBasesTransformer(
        transformer.channel_size, transformer.raw_channels_per_state,
        transformer.num_of_controls_channels, transformer.num_previous_states,
        transformer.num_rewards, transformer.policy_dist_count)
    for b in transformer.board_space:
        index = transformer.channel_size * b.channel_id + b.y_idx * transformer.num_rows + b.x_idx
        c_transformer.add_board_base(b.base_indx, index)
    for c in transformer.control_space:
        c_transformer.add_control_base(c.base_indx, c.channel_id, c.value)
    if device is not None and hasattr(c_transformer, 'finalize'):
        c_transformer.finalize(transformer.num_bases, device)
    return c_transformer


class PollerBase(object):
    """Drives an engine object's ``poll`` with the network until the engine has nothing left to evaluate —
    the contract of ``src/ggpzero/util/cppinterface.py:53-153``: the first poll carries empty arrays, every
    later one the predictions for the rows of the previous batch; ``None`` from the engine ends a loop and
    the next loop primes again. Statistics keep the reference's names (``acc_time_polling`` = engine time,
    ``acc_time_prediction`` = network time)."""
    POLL_AGAIN = "poll_again"

    def __init__(self, sm, nn, batch_size=1024, sleep_between_poll=-1):
        self.sm, self.nn = sm, nn
        self.batch_size = batch_size                  # most rows one poll may hand out
        self.sleep_between_poll = sleep_between_poll
        self.poll_last = None                         # predictions waiting to be fed back
        self.reset_stats()

    def _get_poller(self):
        raise NotImplementedError

    def reset_stats(self):
        self.num_predictions_calls = self.total_predictions = 0
        self.acc_time_polling = self.acc_time_prediction = 0

    def _pending_arrays(self, n_heads):
        import numpy as np
        if self.poll_last is None:
            return [np.zeros(0, dtype=np.float32)] * n_heads
        arrays = list(self.poll_last)
        assert len(arrays) == n_heads
        return arrays

    def poll(self, do_stats=False):
        """One engine call + one network call; returns POLL_AGAIN while the engine wants more."""
        import time
        tr = self.nn.gdl_bases_transformer
        arrays = self._pending_arrays(len(tr.policy_dist_count) + 1)
        t_start = time.time()
        flat = self._get_poller().poll(len(arrays[0]), arrays)
        t_polled = time.time()
        if flat is None:
            self.poll_last = None
            return None
        rows = len(flat) // (tr.num_channels * tr.channel_size)
        assert rows <= self.batch_size
        batch = flat.reshape(rows, tr.num_channels, tr.num_cols, tr.num_rows)     # a view, nothing is copied
        self.poll_last = self.nn.get_model().predict_on_batch(batch)
        if do_stats:
            self.num_predictions_calls += 1
            self.total_predictions += rows
            self.acc_time_polling += t_polled - t_start
            self.acc_time_prediction += time.time() - t_polled
        return self.POLL_AGAIN

    def poll_loop(self, cb=None, do_stats=False):
        """Polls until the engine is done; every 100 polls ``cb`` may stop the loop by returning true."""
        import time
        polls = 0
        while self.poll(do_stats=do_stats) == self.POLL_AGAIN:
            polls += 1
            if cb is not None and polls % 100 == 0 and cb():
                return
            if self.sleep_between_poll > 0:
                time.sleep(self.sleep_between_poll)

    def update_nn(self, nn):
        self.nn = nn

    def dump_stats(self):
        for label, v in (("num of prediction calls", self.num_predictions_calls),
                         ("predictions", self.total_predictions),
                         ("acc_time_polling", self.acc_time_polling),
                         ("acc_time_prediction", self.acc_time_prediction)):
            print(label, v)


class PlayPoller(PollerBase):
    """Match play: a ``ggpzero_interface.Player`` behind the poll loop (``cppinterface.py:156-182``); the
    ``player_*`` methods of the engine object are re-exported on the poller."""
    _FORWARDED = ("reset", "apply_move", "move", "get_move", "update_config", "balance_moves", "tree_debug")

    def __init__(self, sm, nn, conf, **engine_kwds):
        import attr
        from .defs import confs
        assert isinstance(conf, confs.PUCTEvaluatorConfig)
        super().__init__(sm, nn, batch_size=conf.batch_size)
        self.c_transformer = create_c_transformer(nn.gdl_bases_transformer)
        self.c_player = ggpzero_interface.Player(sm, self.c_transformer, attr.asdict(conf), **engine_kwds)
        for suffix in self._FORWARDED:
            setattr(self, "player_" + suffix, getattr(self.c_player, "player_" + suffix))

    def _get_poller(self):
        return self.c_player


class Supervisor(PollerBase):
    """Self-play: a ``ggpzero_interface.Supervisor`` behind the poll loop (``cppinterface.py:185-226``)."""

    def __init__(self, sm, nn, batch_size=1024, sleep_between_poll=-1, workers=None, identifier="",
                 **engine_kwds):
        super().__init__(sm, nn, batch_size=batch_size, sleep_between_poll=sleep_between_poll)
        self.c_transformer = create_c_transformer(nn.gdl_bases_transformer)
        self.c_supervisor = ggpzero_interface.Supervisor(sm, self.c_transformer, batch_size, identifier,
                                                         **engine_kwds)
        if workers:
            self.c_supervisor.set_num_workers(workers)

    def _get_poller(self):
        return self.c_supervisor

    def start_self_play(self, conf, num_workers):
        import attr
        from .defs import confs
        assert isinstance(conf, confs.SelfPlayConfig)
        return self.c_supervisor.start_self_play(num_workers, attr.asdict(conf))

    def fetch_samples(self):
        from .defs import datadesc
        return [datadesc.Sample(**d) for d in (self.c_supervisor.fetch_samples() or [])]

    def add_unique_state(self, s):
        """A no-op, as in the reference's wrapper (it returns before doing anything, cppinterface.py:218-219)."""

    def clear_unique_states(self):
        self.c_supervisor.clear_unique_states()
