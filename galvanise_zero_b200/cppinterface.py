"""Python side of the engine boundary — mirrors ``src/ggpzero/util/cppinterface.py``:
``create_c_transformer`` (``:33-50``), ``PollerBase`` (``:53-153``), ``PlayPoller`` (``:156-182``)
and ``Supervisor`` (``:185-226``)."""
from __future__ import annotations

from . import interface as ggpzero_interface


def create_c_transformer(transformer, device=None):
    """Python ``GdlBasesTransformer`` (bases.py) -> engine-side transformer; same feeding order as
    the reference. With ``device`` given the tables are also uploaded (finalised)."""
    c_transformer = ggpzero_interface.GdlBasesTransformer(
        transformer.channel_size, transformer.raw_channels_per_state,
        transformer.num_of_controls_channels, transformer.num_previous_states,
        transformer.num_rewards, transformer.policy_dist_count)
    for b in transformer.board_space:
        index = transformer.channel_size * b.channel_id + b.y_idx * transformer.num_rows + b.x_idx
        c_transformer.add_board_base(b.base_indx, index)
    for c in transformer.control_space:
        c_transformer.add_control_base(c.base_indx, c.channel_id, c.value)
    if device is not None:
        c_transformer.finalize(transformer.num_bases, device)
    return c_transformer


class PollerBase(object):
    """``src/ggpzero/util/cppinterface.py:53-153``: drives an engine object's ``poll`` with the
    network's ``predict_on_batch`` until the engine has nothing left to evaluate."""
    POLL_AGAIN = "poll_again"

    def __init__(self, sm, nn, batch_size=1024, sleep_between_poll=-1):
        self.sm = sm
        self.nn = nn
        self.batch_size = batch_size
        self.sleep_between_poll = sleep_between_poll
        self.poll_last = None
        self.reset_stats()

    def _get_poller(self):
        raise NotImplementedError

    def reset_stats(self):
        self.num_predictions_calls = 0
        self.total_predictions = 0
        self.acc_time_polling = 0
        self.acc_time_prediction = 0

    def poll(self, do_stats=False):
        """POLL_AGAIN is returned to indicate that poll() has to be called again."""
        import time
        import numpy as np
        transformer = self.nn.gdl_bases_transformer
        expect_num_arrays = len(transformer.policy_dist_count) + 1
        if self.poll_last is None:
            dummy = np.zeros(0, dtype=np.float32)
            arrays = [dummy for _ in range(expect_num_arrays)]
        else:
            arrays = list(self.poll_last)
        assert len(arrays) == expect_num_arrays
        s0 = time.time()
        pred_array = self._get_poller().poll(len(arrays[0]), arrays)
        if pred_array is None:
            self.poll_last = None
            return None
        s1 = time.time()
        t = transformer
        num_predictions = len(pred_array) // (t.num_channels * t.channel_size)
        assert num_predictions <= self.batch_size
        pred_array = pred_array.reshape(num_predictions, t.num_channels, t.num_cols, t.num_rows)
        self.poll_last = self.nn.get_model().predict_on_batch(pred_array)
        if do_stats:
            s2 = time.time()
            self.num_predictions_calls += 1
            self.total_predictions += num_predictions
            self.acc_time_polling += s1 - s0
            self.acc_time_prediction += s2 - s1
        return self.POLL_AGAIN

    def poll_loop(self, cb=None, do_stats=False):
        """polls until done; ``cb`` is called every 100 polls and stops the loop when it returns true."""
        import time
        cb_every_n = 100
        count = 1
        while self.poll(do_stats=do_stats) == self.POLL_AGAIN:
            if count % cb_every_n == 0 and cb is not None and cb():
                break
            count += 1
            if self.sleep_between_poll > 0:
                time.sleep(self.sleep_between_poll)

    def update_nn(self, nn):
        self.nn = nn

    def dump_stats(self):
        print("num of prediction calls", self.num_predictions_calls)
        print("predictions", self.total_predictions)
        print("acc_time_polling", self.acc_time_polling)
        print("acc_time_prediction", self.acc_time_prediction)


class PlayPoller(PollerBase):
    """``src/ggpzero/util/cppinterface.py:156-182``."""

    def __init__(self, sm, nn, conf, **engine_kwds):
        import attr
        from .defs import confs
        assert isinstance(conf, confs.PUCTEvaluatorConfig)
        super().__init__(sm, nn, batch_size=conf.batch_size)
        self.c_transformer = create_c_transformer(nn.gdl_bases_transformer)
        self.c_player = ggpzero_interface.Player(sm, self.c_transformer, attr.asdict(conf), **engine_kwds)
        for name in "reset apply_move move get_move update_config balance_moves tree_debug".split():
            name = "player_" + name
            setattr(self, name, getattr(self.c_player, name))

    def _get_poller(self):
        return self.c_player


class Supervisor(PollerBase):
    """``src/ggpzero/util/cppinterface.py:185-226``."""

    def __init__(self, sm, nn, batch_size=1024, sleep_between_poll=-1, workers=None, identifier="",
                 **engine_kwds):
        self.c_transformer = create_c_transformer(nn.gdl_bases_transformer)
        self.c_supervisor = ggpzero_interface.Supervisor(sm, self.c_transformer, batch_size, identifier,
                                                         **engine_kwds)
        if workers:
            self.c_supervisor.set_num_workers(workers)
        super().__init__(sm, nn, batch_size=batch_size, sleep_between_poll=sleep_between_poll)

    def _get_poller(self):
        return self.c_supervisor

    def start_self_play(self, conf, num_workers):
        import attr
        from .defs import confs
        assert isinstance(conf, confs.SelfPlayConfig)
        return self.c_supervisor.start_self_play(num_workers, attr.asdict(conf))

    def fetch_samples(self):
        from .defs import datadesc
        res = self.c_supervisor.fetch_samples()
        if res:
            return [datadesc.Sample(**d) for d in res]
        return []

    def add_unique_state(self, s):
        return      # the reference returns before doing anything as well (cppinterface.py:218-219)

    def clear_unique_states(self):
        self.c_supervisor.clear_unique_states()
