"""Self-play side of the reference's worker, re-hosted on the device
supervisor: ``Worker.configure_self_play`` / ``on_request_samples`` / ``cb_from_superviser``
(``src/ggpzero/distributed/worker.py:138-206``) without the Twisted transport — the caller owns the
connection and exchanges the same payloads (a ``SelfPlayConfig``, a list of base64 states in, a list of
``Sample`` with base64 states out)."""
from __future__ import annotations

import time

from . import cppinterface, gendata


class SelfPlayWorker(object):
    def __init__(self, sm, nn, self_play_batch_size=1024, inline_manager=False, min_num_samples=128,
                 server_poll_time=10.0, identifier='worker', **engine_kwds):
        self.sm = sm
        self.nn = nn
        self.batch_size = self_play_batch_size
        self.inline_manager = inline_manager
        self.min_num_samples = min_num_samples
        self.server_poll_time = server_poll_time
        self.identifier = identifier
        self.engine_kwds = engine_kwds
        self.supervisor = None
        self.samples = []
        self.last_stats = {}

    def configure_self_play(self, self_play_conf):
        """worker.py:138-152: (re)create the supervisor for a self-play configuration."""
        self.supervisor = cppinterface.Supervisor(self.sm, self.nn, batch_size=self.batch_size,
                                                  identifier=self.identifier, **self.engine_kwds)
        self.supervisor.start_self_play(self_play_conf, 0 if self.inline_manager else 1)

    def update_nn(self, nn):
        self.nn = nn
        if self.supervisor is not None:
            self.supervisor.update_nn(nn)
            self.supervisor.clear_unique_states()

    def _cb(self):
        for sample in self.supervisor.fetch_samples():
            self.samples.append(gendata.encode_sample_states(sample))
        if time.time() > self._t_request + self.server_poll_time:
            return True
        return len(self.samples) > self.min_num_samples

    def on_request_samples(self, new_states=()):
        """worker.py:176-206: poll until enough samples are ready; returns them (states base64-packed)."""
        assert self.supervisor is not None, 'configure_self_play() first'
        self._t_request = time.time()
        self.samples = []
        self.supervisor.reset_stats()
        for s in new_states:
            self.supervisor.add_unique_state(s)
        t0 = time.time()
        self.supervisor.poll_loop(do_stats=True, cb=self._cb)
        dt = max(time.time() - t0, 1e-9)
        sup = self.supervisor
        sup.c_supervisor.check_health()       # halted game slots / dropped samples are an error, not a slow decay
        self.last_stats = {'samples': len(self.samples), 'prediction_calls': sup.num_predictions_calls,
                           'predictions': sup.total_predictions, 'acc_time_polling': sup.acc_time_polling,
                           'acc_time_prediction': sup.acc_time_prediction, 'seconds': dt,
                           'predictions_per_second': sup.total_predictions / dt}
        return self.samples
