"""Self-play driver: engine step + network forward on one stream, optionally as one CUDA graph.

The reference's equivalent is the loop ``Supervisor.poll_loop`` -> ``PollerBase.poll``
(``src/ggpzero/util/cppinterface.py:78-144``): C++ poll, reshape, ``predict_on_batch``, feed back.
Here both halves stay on the device: ``PuctEngine.advance`` writes bf16 planes, the PyTorch tower
reads them in place and writes the policy / value buffers the next ``advance`` consumes.
"""
from __future__ import annotations

import torch

from .engine import PuctEngine


def staggered_start_states(sm, n, max_ply, seed=0xB200, slot0=0):
    """n positions for ``PuctEngine.set_start_states``: slot i sits ``hash(slot0 + i) % max_ply`` uniformly random
    plies into a game (device playouts, ``DeviceStateMachine.playout_step``). A long-running self-play job holds
    its games at all plies at once; starting every slot at the initial state instead measures the opening only.
    Slots whose random game ended early keep the initial state. Returns int32 [n, state_words] on the device."""
    idx = torch.arange(slot0, slot0 + n, dtype=torch.int64, device=sm.device)
    target = ((idx * 2654435761 + (seed & 0xffff)) >> 7) % max(1, int(max_ply))
    s = torch.zeros((n, sm.state_words), dtype=torch.int32, device=sm.device)
    legal, term, goals = sm.alloc_views(n)
    term.fill_(1)                      # the first step resets every slot to the initial state (ply 0)
    out = sm.initial_states(n)
    for ply in range(int(max_ply)):
        sm.playout_step(s, legal, term, goals, seed, ply, slot0=slot0)
        take = (target == ply) & (term == 0)
        out = torch.where(take.unsqueeze(1), s, out)
    return out.contiguous()


class SelfPlayRunner(object):
    def __init__(self, engine: PuctEngine, net, sub_iters=1, use_cuda_graph=True, stream=None):
        self.engine = engine
        self.stream = stream
        self.net = net
        self.sub_iters = sub_iters
        self.graph = None
        self.steps_done = 0
        e = engine
        torch.backends.cudnn.benchmark = True
        self._planes_cl = None
        with torch.no_grad():
            # first step: nothing to feed back yet
            e.advance(None, None, sub_iters)
            self._forward()
            if use_cuda_graph:
                self._capture()
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(e.device))

    def _forward(self):
        e = self.engine
        x = e.planes if e.plane_c_pad else e.planes.contiguous(memory_format=torch.channels_last)
        pol, val = self.net(x)
        e.policy.copy_(pol)
        e.value.copy_(val)

    def _step_eager(self):
        e = self.engine
        # the network forward overlaps the step's side-stream work (re-rooting, goal pass); joined after it
        e.advance(e.policy, e.value, self.sub_iters, defer_join=True)
        self._forward()
        e.join()

    def _capture(self):
        e = self.engine
        s = torch.cuda.Stream(device=e.device)
        s.wait_stream(torch.cuda.current_stream(e.device))
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(3):        # warm up cuDNN / allocator on the side stream
                self._step_eager()
        torch.cuda.current_stream(e.device).wait_stream(s)
        torch.cuda.synchronize(e.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self._step_eager()

    def step(self, n=1):
        with torch.no_grad():
            if self.stream is not None:
                with torch.cuda.stream(self.stream):
                    self._step_n(n)
            else:
                self._step_n(n)
        self.steps_done += n

    def _step_n(self, n):
        for _ in range(n):
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step_eager()

    @property
    def kernels_per_step(self):
        """Launches of this repo's own kernels per step: pre, re-root, propnet, node creation, encode (+ the
        forked goal pass: compaction, gather, propnet, scatter, node creation)."""
        return self.engine.kernels_per_step + 4 * (self.sub_iters - 1)


class CompactSelfPlayRunner(object):
    """Engine step + network forward over the rows that WAIT for the network only (``PuctEngine.compact_requests`` /
    ``advance_compact``) — what the reference's poll hands its network (``src/cpp/supervisor.cpp:101-168``). In a
    long-running job a quarter of the games need no evaluation after a step (their playouts ended in terminal nodes,
    their searches are over). Shapes must be static for CUDA graphs, so one graph is captured per batch size in steps
    of G/32 rows; the host picks the graph from the count of waiting rows the device reported a few steps ago (an
    asynchronous copy into pinned memory, never waited for). A count that grew in the meantime is harmless: requests
    beyond the batch stay pending and their games wait one more step."""

    def __init__(self, engine: PuctEngine, net, sub_iters=1, stream=None, margin=1.03):
        self.engine = e = engine
        self.net = net
        self.sub_iters = sub_iters
        self.stream = stream
        self.margin = margin
        self.steps_done = 0
        self.rows_done = 0
        G = e.n_games
        q = max(1, G // 32)
        self.buckets = sorted({min(G, q * k) for k in range(14, 33)})
        self.count = torch.full((1,), G, dtype=torch.int32).pin_memory()
        self.pol_c = torch.zeros_like(e.policy)
        self.val_c = torch.zeros_like(e.value)
        torch.backends.cudnn.benchmark = True
        self.graphs = {}
        with torch.no_grad():
            e.advance(None, None, sub_iters)
            self._batch(G)
            torch.cuda.synchronize(e.device)
            side = torch.cuda.Stream(device=e.device)
            for b in self.buckets:
                side.wait_stream(torch.cuda.current_stream(e.device))
                with torch.cuda.stream(side):
                    for _ in range(3):
                        self._step_eager(b)
                torch.cuda.current_stream(e.device).wait_stream(side)
                torch.cuda.synchronize(e.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_eager(b)
                self.graphs[b] = g
        if self.stream is not None:
            self.stream.wait_stream(torch.cuda.current_stream(e.device))

    def _batch(self, b):
        e = self.engine
        x = e.compact_requests(b, self.count)
        if not e.plane_c_pad:
            x = x.contiguous(memory_format=torch.channels_last)
        pol, val = self.net(x)
        self.pol_c[:b].copy_(pol)
        self.val_c[:b].copy_(val)

    def _step_eager(self, b):
        e = self.engine
        e.advance_compact(self.pol_c, self.val_c, self.sub_iters, defer_join=True)
        self._batch(b)
        e.join()

    def _pick(self):
        want = int(int(self.count[0]) * self.margin) + 16
        for b in self.buckets:
            if b >= want:
                return b
        return self.buckets[-1]

    def step(self, n=1):
        with torch.no_grad():
            ctx = torch.cuda.stream(self.stream) if self.stream is not None else None
            if ctx is not None:
                ctx.__enter__()
            try:
                for _ in range(n):
                    b = self._pick()
                    self.graphs[b].replay()
                    self.rows_done += b
            finally:
                if ctx is not None:
                    ctx.__exit__(None, None, None)
        self.steps_done += n

    @property
    def kernels_per_step(self):
        """own launches per step: the engine's, plus compaction (list, cap, gather) and the scatter of the predictions"""
        return self.engine.kernels_per_step + 4 + 4 * (self.sub_iters - 1)


class SelfPlayPool(object):
    """One engine+network pipeline per engine on its own CUDA stream, each with its own graph. Used with a
    single engine; several shards go through StaggeredSelfPlay (engine graphs on their own streams, network
    graphs serialised on one), which is what overlaps tree kernels with the tower — the device-side analogue
    of the reference's two managers per worker thread (``src/cpp/supervisor.cpp:79-99,196-231``)."""

    def __init__(self, engines, net, sub_iters=1, use_cuda_graph=True, compact=False):
        if len(engines) != 1:
            raise ValueError('SelfPlayPool drives one engine; use StaggeredSelfPlay for several shards')
        dev = engines[0].device
        self.engines = engines
        self.streams = [torch.cuda.Stream(device=dev) for _ in engines]
        if compact and use_cuda_graph:
            self.runners = [CompactSelfPlayRunner(e, net, sub_iters, stream=s) for e, s in zip(engines, self.streams)]
        else:
            self.runners = [SelfPlayRunner(e, net, sub_iters, use_cuda_graph, stream=s)
                            for e, s in zip(engines, self.streams)]
        self.device = dev

    def step(self, n=1):
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for _ in range(n):
            for r in self.runners:
                r.step(1)
        for s in self.streams:
            cur.wait_stream(s)

    def totals(self):
        out = None
        for e in self.engines:
            t = e.totals()
            if out is None:
                out = dict(t)
            else:
                for k, v in t.items():
                    out[k] = max(out[k], v) if k.startswith('max_') or k == 'first_error' else out[k] + v
        return out

    @property
    def kernels_per_step(self):
        return sum(r.kernels_per_step for r in self.runners)


class StaggeredSelfPlay(object):
    """Q engine shards whose tree kernels run on their own streams while ALL network forwards are
    serialised on one stream. The tree kernels are latency-bound (pointer chasing) and the tower is
    throughput-bound, so shard q's select/backup/expand hides behind the forwards of the other shards.
    Engine and network parts of every shard are separate CUDA graphs chained by events:

        eng stream q :  wait nn_done[q]  -> G_eng[q] -> record eng_done[q]
        nn stream    :  wait eng_done[q] -> G_nn[q]  -> record nn_done[q]

    This is the device-side counterpart of the reference's double-buffered managers
    (``src/cpp/supervisor.cpp:79-99,196-231``)."""

    def __init__(self, engines, net, sub_iters=1):
        dev = engines[0].device
        self.device = dev
        self.engines = engines
        self.net = net
        self.sub_iters = sub_iters
        torch.backends.cudnn.benchmark = True
        self.eng_streams = [torch.cuda.Stream(device=dev, priority=-1) for _ in engines]   # tree kernels first
        self.nn_stream = torch.cuda.Stream(device=dev)
        self.eng_done = [torch.cuda.Event() for _ in engines]
        self.nn_done = [torch.cuda.Event() for _ in engines]
        self.g_eng, self.g_nn = [], []
        cur = torch.cuda.current_stream(dev)
        with torch.no_grad():
            for e in engines:
                e.advance(None, None, sub_iters)
                self._forward(e)
            torch.cuda.synchronize(dev)
            for q, e in enumerate(engines):
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    for _ in range(3):
                        e.advance(e.policy, e.value, sub_iters)
                        self._forward(e)
                cur.wait_stream(side)
                torch.cuda.synchronize(dev)
                ge = torch.cuda.CUDAGraph()
                with torch.cuda.graph(ge):
                    e.advance(e.policy, e.value, sub_iters)
                gn = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gn):
                    self._forward(e)
                self.g_eng.append(ge)
                self.g_nn.append(gn)
            torch.cuda.synchronize(dev)
            for q in range(len(engines)):
                self.nn_done[q].record(cur)
        self.steps_done = 0

    def _forward(self, e):
        x = e.planes if e.plane_c_pad else e.planes.contiguous(memory_format=torch.channels_last)
        pol, val = self.net(x)
        e.policy.copy_(pol)
        e.value.copy_(val)

    def step(self, n=1):
        cur = torch.cuda.current_stream(self.device)
        for s in self.eng_streams + [self.nn_stream]:
            s.wait_stream(cur)
        Q = len(self.engines)
        for _ in range(n):
            for q in range(Q):
                es = self.eng_streams[q]
                es.wait_event(self.nn_done[q])
                with torch.cuda.stream(es):
                    self.g_eng[q].replay()
                self.eng_done[q].record(es)
                self.nn_stream.wait_event(self.eng_done[q])
                with torch.cuda.stream(self.nn_stream):
                    self.g_nn[q].replay()
                self.nn_done[q].record(self.nn_stream)
        for s in self.eng_streams + [self.nn_stream]:
            cur.wait_stream(s)
        self.steps_done += n

    totals = SelfPlayPool.totals

    @property
    def kernels_per_step(self):
        return sum(e.kernels_per_step + 4 * (self.sub_iters - 1) for e in self.engines)
