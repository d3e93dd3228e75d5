#!/usr/bin/env python
"""bench.py — hot-path throughput of the B200-native galvanise_zero self-play path.

    python bench.py --gpus N --steps K --warmup W [--workload propnet|puct] [--impl reference]

Under torchrun (N>1) every rank owns a disjoint range of game slots (no data-path collective;
"weak" scaling); the timed region is bracketed by barrier + synchronize, timed with CUDA events on
the launching stream, max over ranks. Rank 0 prints ONE JSON line.

Workloads
  propnet  BASELINE config "Breakthrough 8x8 propnet: batched uniform-random playouts": one step =
           one playout ply over all slots (one launch of the fused choose+transition kernel).
  puct     BASELINE config "Reversi 8x8 PUCT self-play, 800 evals/move, small residual net,
           4096 concurrent games per GPU" (see galvanise_zero_b200/selfplay.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d.get('bf16_tflops'),
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained'), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0,
            'source': 'fallback'}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(',')])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[1])); mx.append(float(s[2]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                'sw_power_cap'), s[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': max(mx), 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def all_max(x, world, device):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(x, world, device):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ------------------------------------------------------------------------------------------------
PROPNET_SPEC_NCU_TRAFFIC_BYTES = 429494400     # 178.7 MB read + 250.8 MB written
PROPNET_TABLE_NCU_TRAFFIC_BYTES = 284900000    # 170.0 MB read + 114.9 MB written
PROPNET_GAME = 'breakthrough'
PROPNET_SLOTS = 1 << 21          # per GPU; working set 166 MB > 126 MB L2


def transition_bytes(sm):
    """Algorithmic bytes of one transition: state in + out, legal masks out,
    move indices in (int16), terminal, goals."""
    nb = (sm.n_bases + 7) // 8
    return 2 * nb + (sum(sm.role_sizes) + 7) // 8 + 2 * sm.n_roles + 1 + sm.n_roles


def cpu_baseline_propnet(game, seconds_target=12.0, threads=None):
    from galvanise_zero_b200 import tables
    from oracle import pnref
    threads = threads or os.cpu_count() or 1
    blob = tables.blob_by_name(game)
    # calibrate
    secs, tr, _ = pnref.playout_bench(blob, 64 * threads, 50, 0xB200, threads)
    rate = tr / max(secs, 1e-9)
    steps = 200
    slots = max(threads, int(rate * seconds_target / steps))
    secs, tr, _ = pnref.playout_bench(blob, slots, steps, 0xB200, threads)
    return {'value': tr / secs, 'unit': 'transitions/s', 'cores': threads, 'kind': 'port',
            'sample': '%s: %d slots x %d playout plies, %d host threads, %.1f s '
                      '(oracle/propnet_ref.c, nextState+updateBases per ply)' % (
                          game, slots, steps, threads, secs)}


def run_propnet(args, rank, world, local):
    import numpy as np
    import torch
    from galvanise_zero_b200.statemachine import DeviceStateMachine
    from galvanise_zero_b200 import lib as gzlib

    peaks = load_peaks()
    game = args.game or PROPNET_GAME
    dev = DeviceStateMachine(game, device=local, goals_mode=gzlib.GOALS_ALWAYS)
    n = args.slots or PROPNET_SLOTS
    slot0 = rank * n
    s = torch.zeros((n, dev.state_words), dtype=torch.int32, device=dev.device)
    l, t, g = dev.alloc_views(n)
    t.fill_(1)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev.device)
    seed = 0xB200
    step = 0
    # bring the slots to mixed mid-game positions, then warm up
    for _ in range(30 + args.warmup):
        dev.playout_step(s, l, t, g, seed, step, slot0=slot0, counter=None)
        step += 1
    barrier(world)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    cnt.zero_()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    ev0.record()
    for _ in range(args.steps):
        dev.playout_step(s, l, t, g, seed, step, slot0=slot0, counter=cnt)
        step += 1
    ev1.record()
    barrier(world)
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary() if sampler else None
    ms_max = all_max(ms, world, dev.device)
    transitions = all_sum(float(cnt.item()), world, dev.device)
    value = transitions / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (pn_kernel<PLAYOUT>): algorithmic bytes / launch time
    tb = transition_bytes(dev)
    per_launch_bytes = tb * n
    launch_s = (ms * 1e-3) / args.steps
    achieved = per_launch_bytes / launch_s / 1e9
    spec = dev.specialised and n >= (1 << 18)
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch at 2^21 slots (ncu --set full)
                'traffic': (PROPNET_SPEC_NCU_TRAFFIC_BYTES if spec else PROPNET_TABLE_NCU_TRAFFIC_BYTES) if (n == PROPNET_SLOTS and game == PROPNET_GAME) else None,
                'kernel': 'spec_kernel<GzSpec_%s, PLAYOUT> (game-specialised, thread-sliced)' % game if spec
                          else ('pn_wide_kernel<MODE_PLAYOUT>' if dev.launch_info['smem_bytes'] == 0 else 'pn_kernel<MODE_PLAYOUT> (table-driven)'),
                'bytes_per_unit': tb, 'units_per_launch': n, 'launch_ms': launch_s * 1e3,
                'peak_source': peaks['source'] + ' (HBM copy)',
                'traffic_source': 'profiles/r01b_propnet_spec_kernel_ncu_full.md' if spec
                                  else 'profiles/r01_propnet_playout_ncu_full.md'}

    # ---- e2e: the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    e2e = None
    if rank == 0 or world > 1:
        ne = min(n, 1 << 20, max(1024, (256 << 20) // (dev.state_words * 8 + dev.legal_words * 4)))
        hs = torch.empty((ne, dev.state_words), dtype=torch.int32).pin_memory()
        hm = torch.empty((ne, dev.n_roles), dtype=torch.int16).pin_memory()
        hn = torch.empty((ne, dev.state_words), dtype=torch.int32).pin_memory()
        hl = torch.empty((ne, dev.legal_words), dtype=torch.int32).pin_memory()
        ht = torch.empty((ne,), dtype=torch.uint8).pin_memory()
        hg = torch.empty((ne, dev.n_roles), dtype=torch.uint8).pin_memory()
        # a legal joint move for every (non-terminal) state: take the playout kernel's own choice
        mo = torch.zeros((n, dev.n_roles), dtype=torch.int16, device=dev.device)
        s_prev = s.clone()
        dev.playout_step(s, l, t, g, seed, step, slot0=slot0, moves_out=mo)
        step += 1
        live = (mo[:ne, 0] >= 0)
        mo_c = mo[:ne].clone()
        mo_c[~live] = 0
        hs.copy_(s_prev[:ne]); hm.copy_(mo_c)
        torch.cuda.synchronize()
        out = (hn.numpy().view(np.uint32), hl.numpy().view(np.uint32), ht.numpy(), hg.numpy())
        hs_np, hm_np = hs.numpy().view(np.uint32), hm.numpy()
        for _ in range(2):
            dev.transition_host(hs_np, hm_np, out=out)
        reps = max(3, min(args.steps, 10))
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            dev.transition_host(hs_np, hm_np, out=out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = all_max(dt, world, dev.device)
        e2e_val = all_sum(float(ne * reps), world, dev.device) / dt
        e2e = {'value': e2e_val, 'unit': 'transitions/s',
               'h2d_bytes_per_step': int(hs.numel() * 4 + hm.numel() * 2),
               'd2h_bytes_per_step': int(hn.numel() * 4 + hl.numel() * 4 + ht.numel() + hg.numel()),
               'api': 'gz_sm_transition_host (DeviceStateMachine.transition_host), pinned host buffers',
               'units_per_step': ne}

    # ---- the encoder on the same slots (encode GB/s on random-played positions)
    encoder = None
    try:
        from galvanise_zero_b200 import bases, cppinterface, lookup
        from galvanise_zero_b200.defs import datadesc
        tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))
        ct = cppinterface.create_c_transformer(tr, device=local)
        ne = min(n, 1 << 19)
        out = torch.empty((ne, ct.total_size()), dtype=torch.bfloat16, device=dev.device)
        prev = s[:ne].clone().reshape(ne, 1, dev.state_words)
        for _ in range(3):
            ct.encode(s[:ne], prev, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ct.encode(s[:ne], prev, out=out)
        e1.record()
        torch.cuda.synchronize()
        enc_ms = e0.elapsed_time(e1) / 20
        eb = 2 * ((dev.n_bases + 7) // 8) + 2 * ct.total_size()
        ach = eb * ne / (enc_ms * 1e-3) / 1e9
        encoder = {'kernel': 'encode_kernel<bf16>', 'positions_per_launch': ne, 'launch_ms': enc_ms,
                   'bytes_per_unit': eb, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                   'frac': ach / peaks['hbm_gbs'], 'planes': [tr.num_channels, tr.num_cols, tr.num_rows]}
    except Exception as exc:           # a game without a plane description
        encoder = {'unavailable': str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_propnet(game)

    if rank == 0:
        line = {
            'metric': 'propnet transitions/sec', 'value': value, 'unit': 'transitions/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u32', 'data': 'synthetic',
            'config': {'workload': '%s propnet, batched uniform-random playouts' % (
                           'breakthrough 8x8' if game == 'breakthrough' else game),
                       'slots_per_gpu': n, 'l2': 'working set %.0f MB per GPU > 126 MB L2' % (
                           n * (dev.state_words * 4 + dev.legal_words * 4 + 1 + dev.n_roles) / 1e6),
                       'parallelism': 'games sharded across %d GPU(s), no collective' % world,
                       'launch': dev.launch_info},
            'roofline': roofline, 'encoder_roofline': encoder, 'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': args.steps, 'clocks': clocks,
        }
        print(json.dumps(line), flush=True)



# ------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of one puct_pre_kernel launch in the bench configuration
# (ncu --set full capture, profiles/r01c_puct_pre_kernel_ncu_full.md)
PUCT_PRE_NCU_TRAFFIC_BYTES = 57936896
PUCT_GAME = 'reversi'
PUCT_GAMES_PER_GPU = 4096
PUCT_EVALS_PER_MOVE = 800
PUCT_WORKLOAD = ('reversi 8x8 PUCT self-play, 800 evals/move, small residual net (64f x 5), '
                 '4096 concurrent games per GPU')


def puct_selfplay_conf():
    from galvanise_zero_b200.defs import confs
    return confs.selfplay_config_template(evals_per_move=PUCT_EVALS_PER_MOVE)


def cpu_reference_selfplay(seconds=14.0, warm=3.0, threads=None, batch=32):
    """The reference's own C++ PUCT self-play (oracle/_ref: Supervisor worker mode, N threads x 2
    managers x `batch` games, src/cpp/supervisor.cpp:79-99) on the host cores, network replaced by
    the synthetic hash network so that it measures propnet + PUCT + encode (the reference's
    acc_time_polling notion)."""
    from galvanise_zero_b200 import bases, cppinterface, lookup, tables
    from galvanise_zero_b200.defs import datadesc
    from oracle import gzref
    threads = threads or max(1, (os.cpu_count() or 2) - 1)
    gi = lookup.by_name(PUCT_GAME)
    tr = bases.GdlBasesTransformer(gi, datadesc.generation_for(PUCT_GAME))
    ct = cppinterface.create_c_transformer(tr)
    eng = gzref.RefEngine(tables.blob_by_name(PUCT_GAME), ct)
    st = eng.selfplay(puct_selfplay_conf(), batch, threads, max_seconds=seconds, warm_seconds=warm)
    w = max(st['window_seconds'], 1e-9)
    return {'value': st['window_moves'] / w, 'unit': 'moves/s', 'cores': threads + 1, 'kind': 'reference',
            'evals_per_s': st['window_evals'] / w,
            'sample': '%s, reference C++ (puct/evaluator.cpp, selfplay.cpp, supervisor.cpp compiled unmodified '
                      'against shims) in worker mode: %d worker threads x 2 managers x %d games, %d evals/move, '
                      'synthetic network, %.1f s window after %.1f s warm-up'
                      % (PUCT_GAME, threads, batch, PUCT_EVALS_PER_MOVE, w, warm)}


def time_region(fn, device):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1)


def run_puct(args, rank, world, local):
    import numpy as np
    import torch
    from galvanise_zero_b200.engine import PuctEngine
    from galvanise_zero_b200.nn import make_inference_network
    from galvanise_zero_b200.selfplay import SelfPlayPool, StaggeredSelfPlay
    from galvanise_zero_b200 import distributed as gzdist

    peaks = load_peaks()
    torch.cuda.set_device(local)
    G = args.slots or PUCT_GAMES_PER_GPU
    sp = puct_selfplay_conf()
    P = max(1, args.pipelines)
    Gp = G // P
    engines = [PuctEngine(PUCT_GAME, sp, n_games=G // P, mode=0, device=local,
                          seed=gzdist.rank_seed(0xB200, rank) + 104729 * k, node_cap=4096, child_cap=4096 * 16,
                          plane_c_pad=16)
               for k in range(P)]
    eng = engines[0]
    net = make_inference_network(eng.transformer, 'small', seed=1234, device=eng.device)
    if P > 1:
        # several shards: engine graphs on their own streams, network graphs serialised on one
        runner = StaggeredSelfPlay(engines, net, sub_iters=args.sub_iters)
    else:
        runner = SelfPlayPool(engines, net, sub_iters=args.sub_iters, use_cuda_graph=not args.no_graph)
    # bring the games to a desynchronised steady state (untimed), then warm up
    runner.step(args.presteps)
    runner.step(args.warmup)
    barrier(world)
    t_before = runner.totals()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    ev0.record()
    runner.step(args.steps)
    ev1.record()
    barrier(world)
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary() if sampler else None
    t_after = runner.totals()
    d_moves = t_after['moves'] - t_before['moves']
    d_evals = t_after['evals'] - t_before['evals']
    ms_max = all_max(ms, world, eng.device)
    moves = all_sum(d_moves, world, eng.device)
    evals = all_sum(d_evals, world, eng.device)
    value = moves / (ms_max * 1e-3)

    # ---- per-kernel timing, eager, CUDA events on the launching stream: 40 engine steps with events
    # between the launches (gz_puct_advance_timed), each followed by the real network forward so that the
    # trees keep evolving; the same steps give the algorithmic bytes from the engine's traffic counters
    parts = {}
    with torch.no_grad():
        reps = 40
        acc = {}
        t0c = eng.totals()
        pol, val = net(eng.planes)
        for _ in range(reps):
            planes, ms_k = eng.advance_timed(pol, val)
            for k, v in ms_k.items():
                acc[k] = acc.get(k, 0.0) + v
            pol, val = net(planes)
        torch.cuda.synchronize(eng.device)
        t1c = eng.totals()
        for k, v in acc.items():
            parts[k + '_ms'] = v / reps
        parts['engine_advance_ms'] = sum(acc.values()) / reps
        x = eng.planes
        side = torch.cuda.Stream(device=eng.device)
        side.wait_stream(torch.cuda.current_stream(eng.device))
        with torch.cuda.stream(side):
            for _ in range(3):
                net(x)
        torch.cuda.current_stream(eng.device).wait_stream(side)
        torch.cuda.synchronize(eng.device)
        g_nn = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_nn):
            net(x)
        g_nn.replay()
        parts['network_forward_ms'] = time_region(lambda: [g_nn.replay() for _ in range(10)], eng.device) / 10
    # roofline of the dominant hand-written kernel, puct_pre_kernel (select + reply + backup + controller):
    # algorithmic bytes per evaluation = sum over path nodes of (32 B node record +
    # 24 B per child scanned) + 24 B per path node written by backup + 20 B per child created
    d_nodes = t1c['select_nodes'] - t0c['select_nodes']
    d_children = t1c['select_children'] - t0c['select_children']
    d_new = t1c['new_children'] - t0c['new_children']
    d_ev = max(t1c['evals'] - t0c['evals'], 1.0)
    alg_bytes = 32.0 * d_nodes + 24.0 * d_children + 24.0 * d_nodes + 20.0 * d_new
    per_launch = alg_bytes / reps
    pre_s = parts['puct_pre_kernel_ms'] * 1e-3
    achieved = per_launch / pre_s / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'], 'traffic': PUCT_PRE_NCU_TRAFFIC_BYTES,
                'kernel': 'puct_pre_kernel (one warp per game tree: reply + backup + controller + select; '
                          'dependent pointer chasing over a %.1f GB arena, latency-bound)' % (eng.arena_bytes / 1e9),
                'bytes_per_unit': alg_bytes / d_ev, 'unit_name': 'evaluation', 'units_per_launch': d_ev / reps,
                'algorithmic_bytes_per_launch': per_launch,
                'mean_path_nodes_per_eval': d_nodes / d_ev, 'mean_children_scanned_per_eval': d_children / d_ev,
                'mean_children_created_per_eval': d_new / d_ev,
                'launch_ms': parts['puct_pre_kernel_ms'], 'peak_source': peaks['source'] + ' (HBM copy)',
                'traffic_source': 'ncu --set full, profiles/r01c_puct_pre_kernel_ncu_full.md (dram read+write per launch)'}
    flops = net.flops_per_position * Gp
    nn_s = parts['network_forward_ms'] * 1e-3
    nn_ach = flops / nn_s / 1e12
    peak = peaks['bf16_tflops_sustained'] or peaks['bf16_tflops']
    network_roofline = {'bound': 'tensor', 'achieved': nn_ach, 'peak': peak, 'unit': 'TFLOP/s',
                        'frac': nn_ach / peak, 'traffic': None,
                        'kernel': 'residual tower forward (PyTorch module, cuDNN runtime-fused conv+bias+ReLU, bf16 '
                                  'channels_last) — the consumer of the planes, not part of the hot path; %d positions '
                                  'per launch' % Gp,
                        'flops_per_unit': net.flops_per_position, 'units_per_launch': Gp,
                        'peak_source': peaks['source'] + ' (sustained bf16)'}

    # ---- e2e: the reference-shaped numpy poll() API with host buffers (D2H planes, H2D planes for the
    # network as Keras predict_on_batch does, D2H policy/value, H2D into the engine), every step
    e2e = None
    if (rank == 0 or world > 1) and not args.no_e2e:
        from galvanise_zero_b200 import interface
        import attr
        # worker mode = the reference's double-buffered managers (supervisor.cpp:79-99): two managers of
        # batch_size games each; the engine step of one overlaps the network call on the other's batch.
        # Supervisor(batch_size = G/2) + one worker therefore holds the configured G concurrent games.
        inline = args.e2e_inline
        sup = interface.Supervisor(PUCT_GAME, eng.c_transformer, G if inline else G // 2, 'bench', device=local,
                                   seed=0xE2E + rank, node_cap=2048, child_cap=2048 * 16)
        sup.start_self_play(0 if inline else 1, attr.asdict(sp))
        R = len(eng.role_sizes)
        h2d = d2h = 0
        Gm = sup.batch_size                                     # rows per poll

        # the user's network call (the reference's model.predict_on_batch): host planes in, host
        # predictions out (one array per policy head + value, each in its own pinned buffer); the forward
        # itself is a CUDA graph over a fixed-size device batch
        x_dev = torch.zeros((Gm,) + eng.plane_shape, dtype=torch.float32, device=eng.device)
        out_dev = [torch.zeros((Gm, w), dtype=torch.float32, device=eng.device)
                   for w in list(eng.role_sizes) + [eng.n_rewards]]
        out_pin = [torch.empty((Gm, w), dtype=torch.float32).pin_memory() for w in list(eng.role_sizes) + [eng.n_rewards]]
        out_np = [t.numpy() for t in out_pin]

        def nn_static():
            xb = x_dev.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            p_, v_ = net(xb)
            off = 0
            for r in range(R):
                out_dev[r].copy_(p_[:, off:off + eng.role_sizes[r]])
                off += eng.role_sizes[r]
            out_dev[R].copy_(v_)
        with torch.no_grad():
            side2 = torch.cuda.Stream(device=eng.device)
            side2.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(side2):
                for _ in range(3):
                    nn_static()
            torch.cuda.current_stream(eng.device).wait_stream(side2)
            torch.cuda.synchronize(eng.device)
            g_e2e = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_e2e):
                nn_static()

        t_poll = t_nn = 0.0

        def e2e_step(pred):
            nonlocal h2d, d2h, t_poll, t_nn
            n, arrays = pred
            tp0 = time.perf_counter()
            planes = sup.poll(n, arrays)                 # numpy view of pinned engine memory
            tp1 = time.perf_counter()
            t_poll += tp1 - tp0
            if planes is None:
                return (0, [])
            n_req = planes.size // eng.total_size
            d2h += planes.nbytes
            xin = torch.from_numpy(planes).view((n_req,) + eng.plane_shape)
            x_dev[:n_req].copy_(xin, non_blocking=True)
            h2d += planes.nbytes
            g_e2e.replay()
            for k in range(R + 1):
                out_pin[k][:n_req].copy_(out_dev[k][:n_req], non_blocking=True)
            torch.cuda.current_stream(eng.device).synchronize()
            arrs = [a[:n_req] for a in out_np]           # C-contiguous views, no host copies
            nb = sum(a.nbytes for a in arrs)
            d2h += nb
            h2d += nb
            t_nn += time.perf_counter() - tp1
            return (n_req, arrs)

        pred = (0, [])
        for _ in range(min(args.presteps, 2000) * len(sup.managers)):     # games desynchronise before timing
            pred = e2e_step(pred)
        m0 = sup.totals()
        h2d = d2h = 0
        t_poll = t_nn = 0.0
        reps = max(20, min(args.steps, 200)) * len(sup.managers)
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            pred = e2e_step(pred)
        torch.cuda.synchronize()
        dt = all_max(time.perf_counter() - t0, world, eng.device)
        m1 = sup.totals()
        e2e_moves = all_sum(m1['moves'] - m0['moves'], world, eng.device)
        e2e_evals = all_sum(m1['evals'] - m0['evals'], world, eng.device)
        e2e = {'value': e2e_moves / dt, 'unit': 'moves/s', 'h2d_bytes_per_step': int(h2d / reps) * len(sup.managers),
               'd2h_bytes_per_step': int(d2h / reps) * len(sup.managers), 'evals_per_s': e2e_evals / dt,
               'api': 'interface.Supervisor.poll (numpy in / numpy out, reference contract), network fed from pinned host memory',
               'managers': '%d x %d games (%s)' % (len(sup.managers), Gm, 'inline' if len(sup.managers) == 1 else
                                                   'worker mode: engine step of one manager overlaps the network call of the other'),
               'polls': reps, 'rows_per_poll': Gm, 'host_ms_per_poll': {'poll': t_poll * 1e3 / reps, 'network_call': t_nn * 1e3 / reps}}

    # ---- end of 'generation': gather the finished samples to rank 0 (the only collective; untimed)
    recs = np.concatenate([e.fetch_sample_records() for e in engines], axis=0)
    gathered = gzdist.gather_sample_records(torch.from_numpy(recs).to(eng.device), dst=0)
    n_gathered = int(gathered.shape[0]) if gathered is not None else 0

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_selfplay()
        except Exception as exc:   # oracle/_ref missing on this box
            cpu = {'value': None, 'unit': 'moves/s', 'cores': os.cpu_count(), 'kind': 'reference',
                   'sample': 'unavailable: %s' % exc}

    if rank == 0:
        line = {
            'metric': 'PUCT self-play moves/sec', 'value': value, 'unit': 'moves/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': PUCT_WORKLOAD, 'games_per_gpu': G, 'evals_per_move': PUCT_EVALS_PER_MOVE,
                       'network': 'random-init small tower, %.1f MFLOP/position' % (net.flops_per_position / 1e6),
                       'step': 'one engine advance (select/backup + propnet + expand + encode) + one network '
                               'forward over %d leaves' % G,
                       'l2': 'tree arenas %.1f GB per GPU >> 126 MB L2' % (sum(e.arena_bytes for e in engines) / 1e9),
                       'presteps': args.presteps, 'cuda_graph': not args.no_graph,
                       'pipelines': '%d engine+network pipelines of %d games each on separate streams' % (P, G // P),
                       'parallelism': 'games sharded across %d GPU(s), no collective on the hot path' % world},
            'evals_per_s': evals / (ms_max * 1e-3), 'moves_in_window': moves,
            'roofline': roofline, 'network_roofline': network_roofline, 'kernel_times_ms': parts,
            'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': args.steps * runner.kernels_per_step, 'clocks': clocks,
            'engine_totals': {k: t_after[k] for k in ('errors', 'halted', 'max_nodes', 'max_children', 'games', 'samples')},
            'samples_gathered_to_rank0': n_gathered,
        }
        print(json.dumps(line), flush=True)


def run_reference_puct(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step = 6.0
    total = max(1, args.steps) * per_step
    cpu = cpu_reference_selfplay(seconds=min(total, 150.0) + 3.0, warm=3.0)
    dt = time.perf_counter() - t0
    print(json.dumps({
        'impl': 'reference', 'metric': 'PUCT self-play moves/sec', 'value': cpu['value'], 'unit': 'moves/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3 / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': PUCT_WORKLOAD, 'evals_per_move': PUCT_EVALS_PER_MOVE,
                   'note': 'host-core run of the reference C++ search; concurrent games bounded to what the '
                           'host threads hold (the GPU arm holds 4096 per GPU)'},
        'evals_per_s': cpu['evals_per_s'], 'cpu_baseline': cpu,
        'e2e': {'value': cpu['value'], 'unit': 'moves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}),
        flush=True)


def run_reference(args):
    """CPU reference arm: the oracle port of the reference's state-machine loop on all host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    from galvanise_zero_b200 import tables
    from oracle import pnref
    blob = tables.blob_by_name(PROPNET_GAME)
    secs, tr, _ = pnref.playout_bench(blob, 64 * threads, 50, 0xB200, threads)
    rate = tr / max(secs, 1e-9)
    plies = 100
    slots = max(threads, int(rate * 3.0 / plies))       # ~3 s of CPU work per step
    for _ in range(args.warmup):
        pnref.playout_bench(blob, max(threads, slots // 8), plies, 0xB200, threads)
    t0 = time.perf_counter()
    total = 0
    for k in range(args.steps):
        secs, tr, _ = pnref.playout_bench(blob, slots, plies, 0xB200 + k, threads)
        total += tr
    dt = time.perf_counter() - t0
    val = total / dt
    cpu = {'value': val, 'unit': 'transitions/s', 'cores': threads, 'kind': 'port',
           'sample': '%s: %d slots x %d plies per step, %d host threads (oracle/propnet_ref.c; ggplib '
                     'is absent so the reference state machine itself cannot be built)' % (
                         PROPNET_GAME, slots, plies, threads)}
    print(json.dumps({
        'impl': 'reference', 'metric': 'propnet transitions/sec', 'value': val,
        'unit': 'transitions/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3 / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': 'breakthrough 8x8 propnet, batched uniform-random playouts'},
        'cpu_baseline': cpu,
        'e2e': {'value': val, 'unit': 'transitions/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=1000)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='puct', choices=['propnet', 'puct'])
    ap.add_argument('--sub-iters', type=int, default=1)
    ap.add_argument('--presteps', type=int, default=3000)
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--pipelines', type=int, default=1)
    ap.add_argument('--staggered', action='store_true', help='(implied by --pipelines > 1)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-inline', action='store_true')
    ap.add_argument('--slots', type=int, default=0)
    ap.add_argument('--game', default='', help='propnet workload: another compiled game (default breakthrough)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        if args.workload == 'puct':
            run_reference_puct(args)
        else:
            run_reference(args)
        return
    rank, world, local = dist_setup(args.gpus)
    if args.workload == 'propnet':
        run_propnet(args, rank, world, local)
    else:
        run_puct(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
