"""CPU: the poll protocol of cppinterface.PollerBase (src/ggpzero/util/cppinterface.py:78-144) against a
scripted engine object — priming poll with empty arrays, prediction count fed back, None ends the loop,
the callback cadence — and the reference helper buf_to_tuple_reverse_bytes."""
import numpy as np

from galvanise_zero_b200 import cppinterface


class _Transformer(object):
    policy_dist_count = [5, 7]
    num_channels, num_cols, num_rows = 3, 4, 2
    channel_size = 8
    num_rewards = 2


class _Model(object):
    def __init__(self):
        self.shapes = []

    def predict_on_batch(self, x):
        self.shapes.append(x.shape)
        n = x.shape[0]
        return [np.zeros((n, 5), np.float32), np.zeros((n, 7), np.float32), np.zeros((n, 2), np.float32)]


class _NN(object):
    gdl_bases_transformer = _Transformer()

    def __init__(self):
        self.model = _Model()

    def get_model(self):
        return self.model


class _Engine(object):
    def __init__(self, script):
        self.script = list(script)
        self.calls = []

    def poll(self, n, arrays):
        self.calls.append((n, [a.shape for a in arrays]))
        k = self.script.pop(0)
        if k is None:
            return None
        return np.zeros(k * 3 * 8, dtype=np.float32)


class _Poller(cppinterface.PollerBase):
    def __init__(self, engine, nn):
        super().__init__(None, nn, batch_size=16)
        self.engine = engine

    def _get_poller(self):
        return self.engine


def test_poll_protocol():
    eng = _Engine([4, 16, 1, None, 2, None])
    nn = _NN()
    p = _Poller(eng, nn)
    p.poll_loop(do_stats=True)
    assert [c[0] for c in eng.calls] == [0, 4, 16, 1]
    assert eng.calls[0][1] == [(0,), (0,), (0,)]                 # priming call: empty arrays
    assert eng.calls[1][1] == [(4, 5), (4, 7), (4, 2)]
    assert nn.model.shapes == [(4, 3, 4, 2), (16, 3, 4, 2), (1, 3, 4, 2)]
    assert p.poll_last is None and p.total_predictions == 21 and p.num_predictions_calls == 3
    # re-entrant after None: primes again
    p.poll_loop()
    assert [c[0] for c in eng.calls][4:] == [0, 2]


def test_poll_loop_callback_stops():
    eng = _Engine([1] * 1000)
    p = _Poller(eng, _NN())
    seen = []
    p.poll_loop(cb=lambda: seen.append(len(eng.calls)) or len(seen) >= 2)
    assert seen == [100, 200]


def test_buf_to_tuple_reverse_bytes():
    from galvanise_zero_b200.interface import buf_to_tuple_reverse_bytes
    bits = np.random.RandomState(0).randint(0, 2, 77).astype(np.uint8)
    packed = np.packbits(bits).tobytes()
    out = buf_to_tuple_reverse_bytes(packed)
    assert out[:77] == tuple(bool(b) for b in bits) and not any(out[77:])
