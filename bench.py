#!/usr/bin/env python
"""bench.py — hot-path throughput of the B200-native galvanise_zero self-play path.

    python bench.py --gpus N --steps K --warmup W [--workload propnet|puct] [--impl reference]

Under torchrun (N>1) every rank owns a disjoint range of game slots (no data-path collective;
"weak" scaling); the timed region is bracketed by barrier + synchronize, timed with CUDA events on
the launching stream, max over ranks. Rank 0 prints ONE JSON line.

Workloads
  puct     (default) BASELINE config "Reversi 8x8 PUCT self-play, 800 evals/move, small residual net,
           4096 concurrent games per GPU". One bench step = --advances-per-step (256) engine advances,
           each followed by one network forward over the leaves, so that the timed region lasts seconds.
           Every slot's first game starts a uniformly random number of random plies into a game and the
           run is warmed up (--presteps advances), so the games sit at all plies at once as they do in
           a long-running job; the reference arm and the in-line CPU baseline get the same treatment.
           The line also carries BASELINE's first metric under "propnet" (the propnet workload below,
           measured in the same run on the same GPUs).
  propnet  BASELINE config "Breakthrough 8x8 propnet: batched uniform-random playouts": one step =
           one playout ply over all slots (one launch of the fused choose+transition kernel).
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


def run_propnet(args, rank, world, local, steps=None, cpu_seconds=12.0):
    """Returns the propnet line (rank 0) or None. `steps` = playout launches in the timed region."""
    import numpy as np
    import torch
    from galvanise_zero_b200.statemachine import DeviceStateMachine
    from galvanise_zero_b200 import lib as gzlib

    peaks = load_peaks()
    steps = steps or args.steps
    game = (args.game if args.workload == 'propnet' else '') or PROPNET_GAME
    # goal values for terminal states only: what the reference's path reads (src/cpp/puct/node.cpp:153-221; its
    # random playouts, src/cpp/pyobjects/gdltransformer_impl.cpp:50-124, read none at all)
    dev = DeviceStateMachine(game, device=local, goals_mode=gzlib.GOALS_TERMINAL_ONLY)
    if args.group_min_batch is not None:
        dev.set_group_min_batch(args.group_min_batch)
    n = (args.slots if args.workload == 'propnet' else 0) or PROPNET_SLOTS
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
    for _ in range(steps):
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
    launch_s = (ms * 1e-3) / steps
    achieved = per_launch_bytes / launch_s / 1e9
    spec = dev.specialised and n >= (1 << 18)
    ps = dev.program_stats
    wide = ps['wide']
    # the group kernel is the default for wide-format games from 1024 states up; others only when asked
    group_used = (not spec) and ((args.group_min_batch is None and wide and n >= 1024) or
                                 (args.group_min_batch is not None and 0 <= args.group_min_batch <= n))
    # ---- the secondary bound: gate work. A table-driven kernel evaluates the next-state cone and
    # the legal / terminal cone once per 32-game tile (the goal cone only for tiles holding a terminal state): one
    # warp instruction handles 32 literals (lane = gate) of 32 games (bit-sliced word), and a literal costs about 3
    # of them (program word, value word, combine) plus ~10 per 32-gate chunk. The specialised kernel's count is
    # measured (ncu, profiles/r01b_propnet_spec_kernel_ncu_full.md: 66.9 M warp instructions per 2^21 transitions).
    SM_COUNT, SCHED_PER_SM, SM_HZ = 148, 4, 1.965e9
    issue_rate = SM_COUNT * SCHED_PER_SM * SM_HZ                       # warp instructions / s, whole GPU
    lits = ps['literals'][0] + ps['literals'][1]
    chunks = ps['chunks'][0] + ps['chunks'][1]
    if spec:
        wi_per_transition, wi_src = 66.9e6 / (1 << 21), 'ncu smsp__inst_executed.sum of the specialised kernel'
    else:
        wi_per_transition, wi_src = (3.0 * lits / 32 + 10.0 * chunks) / 32, 'model: 3 warp instructions per 32 literals + 10 per chunk, per 32-game tile'
    issue_bound = issue_rate / wi_per_transition
    compute = {'bound': 'issue', 'gates': ps['gates'], 'literals': ps['literals'], 'chunks': ps['chunks'],
               'warp_instructions_per_transition': wi_per_transition, 'source': wi_src,
               'peak': issue_bound, 'unit': 'transitions/s', 'achieved': value, 'frac': value / (issue_bound * world),
               'note': 'next-state + legal/terminal cones (the goal cone runs only for tiles with a terminal state); '
                       'peak = %d SMs x %d schedulers x %.3f GHz / warp instructions per transition, per GPU' % (SM_COUNT, SCHED_PER_SM, SM_HZ / 1e9)}
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': achieved / peaks['hbm_gbs'],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch at 2^21 slots (ncu --set full)
                'traffic': (PROPNET_SPEC_NCU_TRAFFIC_BYTES if spec else PROPNET_TABLE_NCU_TRAFFIC_BYTES) if (n == PROPNET_SLOTS and game == PROPNET_GAME) else None,
                'kernel': 'spec_kernel<GzSpec_%s, PLAYOUT> (game-specialised, thread-sliced)' % game if spec
                          else ('pn_group_kernel<MODE_PLAYOUT> (1024 games per CTA / cluster, one tile per lane)'
                                if group_used
                                else ('pn_wide_kernel<MODE_PLAYOUT>' if wide
                                      else 'pn_kernel<MODE_PLAYOUT> (table-driven)')),
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
        reps = 10
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
        col = ne >= 16384 and ct.total_size() % 8 == 0
        encoder = {'kernel': 'encode_col_kernel<bf16> (column-owned, csrc/gz_encoder_col.cuh)' if col
                             else 'encode_fast_kernel<bf16> / encode_kernel<bf16>', 'positions_per_launch': ne, 'launch_ms': enc_ms,
                   'bytes_per_unit': eb, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                   'frac': ach / peaks['hbm_gbs'], 'planes': [tr.num_channels, tr.num_cols, tr.num_rows],
                   # dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full; reversi / breakthrough
                   # planes 5x8x8 at 2^19 positions): no more than the algorithmic bytes
                   'traffic': 310903552 if (col and ne == (1 << 19) and ct.total_size() == 320) else None,
                   'traffic_source': 'profiles/r02_encode_col_kernel_ncu_full.md'}
    except Exception as exc:           # a game without a plane description
        encoder = {'unavailable': str(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_propnet(game, seconds_target=cpu_seconds)
    del s, l, t, g
    torch.cuda.empty_cache()

    if rank == 0:
        line = {
            'metric': 'propnet transitions/sec', 'value': value, 'unit': 'transitions/s',
            'n_gpus': world, 'steps': steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / steps, 'timed_region_s': ms_max * 1e-3,
            'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u32', 'data': 'synthetic',
            'config': {'workload': '%s propnet, batched uniform-random playouts' % (
                           'breakthrough 8x8' if game == 'breakthrough' else game),
                       'slots_per_gpu': n, 'l2': 'working set %.0f MB per GPU > 126 MB L2' % (
                           n * (dev.state_words * 4 + dev.legal_words * 4 + 1 + dev.n_roles) / 1e6),
                       'parallelism': 'games sharded across %d GPU(s), no collective' % world,
                       'launch': dev.launch_info},
            'roofline': roofline, 'compute_roofline': compute, 'encoder_roofline': encoder, 'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': steps, 'clocks': clocks,
        }
        return line
    return None



# ------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of one puct_pre_kernel launch in the bench configuration
# (ncu --set full capture in the bench's own steady-state regime, profiles/r02_puct_pre_kernel_ncu_full.md)
PUCT_PRE_NCU_TRAFFIC_BYTES = 40281600     # 20.21 MB read + 20.07 MB written, final capture of round 2
PUCT_PRE_NCU_TRAFFIC_SOURCE = 'ncu --set full, profiles/r02_puct_pre_kernel_ncu_full.md (dram read+write per launch)'
PUCT_GAME = 'reversi'
PUCT_EVALS_PER_MOVE = 800
# per game: concurrent games per GPU, tree arena sizes, and the ply range the slots' first games are spread over
# (about the length of a uniformly random game, so that the slots sit at all plies at once)
PUCT_PRESETS = {
    'reversi': dict(games=4096, node_cap=4096, child_cap=4096 * 16, stagger_plies=58,
                    workload='reversi 8x8 PUCT self-play, 800 evals/move, small residual net (64f x 5), '
                             '4096 concurrent games per GPU'),
    'breakthrough': dict(games=4096, node_cap=4096, child_cap=4096 * 24, stagger_plies=50,
                         workload='breakthrough 8x8 PUCT self-play, 800 evals/move, small residual net'),
    'hexLG13': dict(games=1024, node_cap=4096, child_cap=4096 * 96, stagger_plies=100,
                    workload='hexLG13 (13x13, swap rule) PUCT self-play, 800 evals/move, small residual net, '
                             '1024 concurrent games per GPU'),
    'connect6': dict(games=1024, node_cap=4096, child_cap=4096 * 128, stagger_plies=60,
                     workload='connect6 19x19 PUCT self-play, 800 evals/move, small residual net'),
    'amazons_10x10': dict(games=512, node_cap=4096, child_cap=4096 * 512, stagger_plies=50,
                          workload='amazons 10x10 PUCT self-play, 800 evals/move, small residual net'),
    'chess_15d': dict(games=1024, node_cap=4096, child_cap=4096 * 32, stagger_plies=60,
                      workload='chess (15-move draw rule) PUCT self-play, 800 evals/move, small residual net, '
                               '1024 concurrent games per GPU'),
}
CPU_BATCH = 8          # games per reference manager (2 managers per worker thread) in the CPU arms
CPU_WARM_S = 2.5       # one CPU sample: untimed warm-up ...
CPU_WINDOW_S = 7.5     # ... and timed window, in a fresh process with its own seeds
CPU_INLINE_SAMPLES = 3     # the in-line cpu_baseline = mean over this many samples (30 s); the reference arm takes
CPU_SEED0 = 0x5EED0000     # one sample per 2 contract steps, from the same seed sequence


def puct_game(args):
    return (args.game if args.workload == 'puct' else '') or PUCT_GAME


def puct_selfplay_conf():
    from galvanise_zero_b200.defs import confs
    return confs.selfplay_config_template(evals_per_move=PUCT_EVALS_PER_MOVE)


def cpu_reference_selfplay_mean(game, n_samples):
    """Mean over `n_samples` bounded samples (fresh process and seeds each) of the reference's self-play rate."""
    runs = [cpu_reference_selfplay(game, rng_seed=CPU_SEED0 + 7919 * i) for i in range(n_samples)]
    moves = sum(r['window_moves'] for r in runs)
    secs = sum(r['window_s'] for r in runs)
    cpu = dict(runs[0])
    cpu['value'] = moves / secs
    cpu['evals_per_s'] = sum(r['evals_per_s'] * r['window_s'] for r in runs) / secs
    cpu['samples'] = [r['value'] for r in runs]
    cpu['window_s'] = secs
    cpu['window_moves'] = moves
    cpu['sample'] = '%d x [%s]' % (n_samples, runs[0]['sample'])
    return cpu


def cpu_reference_selfplay(game=PUCT_GAME, seconds=CPU_WINDOW_S, warm=CPU_WARM_S, threads=None, batch=CPU_BATCH,
                           rng_seed=CPU_SEED0):
    """The reference's own C++ PUCT self-play (oracle/_ref: Supervisor worker mode, N threads x 2
    managers x `batch` games, src/cpp/supervisor.cpp:79-99) on the host cores, network replaced by
    the synthetic hash network so that it measures propnet + PUCT + encode (the reference's
    acc_time_polling notion). Same regime as the GPU arm: every slot's first game starts a uniformly random
    number of random plies into a game, then `warm` seconds untimed, then the window. Runs in a child
    process (the reference's worker threads cannot be stopped, supervisor.cpp:34-40)."""
    threads = threads or max(1, (os.cpu_count() or 2) - 1)
    code = (
        'import json, sys\n'
        'sys.path.insert(0, %r)\n'
        'from galvanise_zero_b200 import bases, cppinterface, lookup, tables\n'
        'from galvanise_zero_b200.defs import datadesc, confs\n'
        'from oracle import gzref\n'
        'game = %r\n'
        'tr = bases.GdlBasesTransformer(lookup.by_name(game), datadesc.generation_for(game))\n'
        'ct = cppinterface.create_c_transformer(tr)\n'
        'eng = gzref.RefEngine(tables.blob_by_name(game), ct)\n'
        'sp = confs.selfplay_config_template(evals_per_move=%d)\n'
        'st = eng.selfplay(sp, %d, %d, max_seconds=%r, warm_seconds=%r, stagger_max_ply=%d, rng_seed=%d)\n'
        'st.pop("records", None)\n'
        'print("GZREF " + json.dumps(st), flush=True)\n'
        'import os; os._exit(0)\n'
    ) % (ROOT, game, PUCT_EVALS_PER_MOVE, batch, threads, seconds + warm, warm,
         PUCT_PRESETS[game]['stagger_plies'], rng_seed)
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=seconds + warm + 240)
    st = None
    for ln in out.stdout.splitlines():
        if ln.startswith('GZREF '):
            st = json.loads(ln[6:])
    if st is None:
        raise RuntimeError('reference self-play failed: %s' % (out.stderr[-400:] or out.stdout[-400:]))
    w = max(st['window_seconds'], 1e-9)
    return {'value': st['window_moves'] / w, 'unit': 'moves/s', 'cores': threads + 1, 'kind': 'reference',
            'evals_per_s': st['window_evals'] / w, 'window_s': w, 'window_moves': st['window_moves'],
            'sample': '%s, reference C++ (puct/evaluator.cpp, selfplay.cpp, supervisor.cpp compiled unmodified '
                      'against shims) in worker mode: %d worker threads x 2 managers x %d games, %d evals/move, '
                      'synthetic network; slots start at staggered plies (uniform in [0, %d)), %.1f s window after '
                      '%.1f s warm-up. The reference search never returns from a fully explored, unconverged '
                      'end-game tree (evaluator.cpp:782-841), which stalls its manager: its rate decays with '
                      'the length of the run, so both CPU arms use this same bounded sample'
                      % (game, threads, batch, PUCT_EVALS_PER_MOVE, PUCT_PRESETS[game]['stagger_plies'], w, warm)}


def time_region(fn, device):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1)


def run_puct(args, rank, world, local, propnet_line=None):
    import numpy as np
    import torch
    from galvanise_zero_b200.engine import PuctEngine
    from galvanise_zero_b200.nn import make_inference_network
    from galvanise_zero_b200.selfplay import SelfPlayPool, StaggeredSelfPlay, staggered_start_states
    from galvanise_zero_b200 import distributed as gzdist

    peaks = load_peaks()
    torch.cuda.set_device(local)
    game = puct_game(args)
    preset = PUCT_PRESETS[game]
    G = args.slots or preset['games']
    sp = puct_selfplay_conf()
    P = max(1, args.pipelines)
    Gp = G // P
    K = max(1, args.advances_per_step)
    engines = [PuctEngine(game, sp, n_games=Gp, mode=0, device=local,
                          seed=gzdist.rank_seed(0xB200, rank) + 104729 * k, node_cap=preset['node_cap'],
                          child_cap=preset['child_cap'], plane_c_pad=16 if game == PUCT_GAME else 0)
               for k in range(P)]
    eng = engines[0]
    if args.group_min_batch is not None:
        eng.sm.set_group_min_batch(args.group_min_batch)
    if args.max_iters or args.max_work:
        for e in engines:
            e.set_step_limits(args.max_iters or 16, args.max_work or (1 << 20))
    # every slot's first game starts a uniformly random number of random plies into a game
    if not args.no_stagger:
        for k, e in enumerate(engines):
            e.set_start_states(staggered_start_states(e.sm, Gp, preset['stagger_plies'], seed=0xB200 + rank,
                                                      slot0=(rank * P + k) * Gp))
    net = make_inference_network(eng.transformer, 'small', seed=1234, device=eng.device)
    if P > 1:
        # several shards: engine graphs on their own streams, network graphs serialised on one
        runner = StaggeredSelfPlay(engines, net, sub_iters=args.sub_iters)
    else:
        runner = SelfPlayPool(engines, net, sub_iters=args.sub_iters, use_cuda_graph=not args.no_graph,
                              compact=not args.full_batches)
    # untimed: let the trees grow and the searches desynchronise, then the contract's warm-up steps
    runner.step(args.presteps)
    runner.step(args.warmup * K)
    barrier(world)
    t_before = runner.totals()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    ev0.record()
    rows0 = sum(getattr(r, 'rows_done', 0) for r in getattr(runner, 'runners', []))
    runner.step(args.steps * K)
    rows1 = sum(getattr(r, 'rows_done', 0) for r in getattr(runner, 'runners', []))
    ev1.record()
    barrier(world)
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.summary() if sampler else None
    t_after = runner.totals()
    d_moves = t_after['moves'] - t_before['moves']
    d_evals = t_after['evals'] - t_before['evals']
    d_playouts = t_after['playouts'] - t_before['playouts']
    ms_max = all_max(ms, world, eng.device)
    moves = all_sum(d_moves, world, eng.device)
    evals = all_sum(d_evals, world, eng.device)
    value = moves / (ms_max * 1e-3)
    advances = args.steps * K
    regime = {'mean_ply_start': t_before['depth_sum'] / G, 'mean_ply_end': t_after['depth_sum'] / G,
              'nn_batch_fill': d_evals / max(1.0, float(advances) * G),
              'playouts_per_eval': d_playouts / max(d_evals, 1.0),
              # compacted batches: rows the network actually ran over per advance (0: full batches of G rows)
              'network_rows_per_advance': (rows1 - rows0) / float(max(1, args.steps * K)),
              'games_finished_in_window': t_after['games'] - t_before['games'],
              'games_finished_total': t_after['games']}

    # ---- per-kernel timing, eager, CUDA events on the launching stream: engine steps with events
    # between the launches (gz_puct_advance_timed), each followed by the real network forward so that the
    # trees keep evolving; the same steps give the algorithmic bytes from the engine's traffic counters
    parts = {}
    with torch.no_grad():
        reps = 200
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
        parts['network_forward_ms'] = time_region(lambda: [g_nn.replay() for _ in range(50)], eng.device) / 50
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
                'frac': achieved / peaks['hbm_gbs'],
                'traffic': PUCT_PRE_NCU_TRAFFIC_BYTES if (game == PUCT_GAME and Gp == 4096) else None,
                'kernel': 'puct_pre_kernel (one warp per game tree: reply + backup + controller + select; '
                          'dependent pointer chasing over a %.1f GB arena, latency-bound)' % (eng.arena_bytes / 1e9),
                'bytes_per_unit': alg_bytes / d_ev, 'unit_name': 'evaluation', 'units_per_launch': d_ev / reps,
                'algorithmic_bytes_per_launch': per_launch,
                'mean_path_nodes_per_eval': d_nodes / d_ev, 'mean_children_scanned_per_eval': d_children / d_ev,
                'mean_children_created_per_eval': d_new / d_ev,
                'launch_ms': parts['puct_pre_kernel_ms'], 'peak_source': peaks['source'] + ' (HBM copy)',
                'traffic_source': PUCT_PRE_NCU_TRAFFIC_SOURCE}
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
        sup = interface.Supervisor(game, eng.c_transformer, G if inline else G // 2, 'bench', device=local,
                                   seed=0xE2E + rank, node_cap=max(1024, preset['node_cap'] // 2),
                                   child_cap=max(1024, preset['child_cap'] // 2))
        if not args.no_stagger:
            sup.start_states_fn = lambda e, k: staggered_start_states(
                e.sm, e.n_games, preset['stagger_plies'], seed=0xE2E + rank, slot0=(2 * rank + k) * e.n_games)
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

        # poll hands out only the rows that wait for the network (the reference's contract): the forward runs over the
        # smallest captured batch size that holds them (steps of Gm/8 rows), not over all Gm rows
        buckets = sorted({max(1, (Gm * k) // 8) for k in range(2, 9)})

        def nn_static(b):
            xb = x_dev[:b].to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
            p_, v_ = net(xb)
            off = 0
            for r in range(R):
                out_dev[r][:b].copy_(p_[:, off:off + eng.role_sizes[r]])
                off += eng.role_sizes[r]
            out_dev[R][:b].copy_(v_)
        g_e2e = {}
        with torch.no_grad():
            side2 = torch.cuda.Stream(device=eng.device)
            side2.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(side2):
                for b in buckets:
                    for _ in range(3):
                        nn_static(b)
            torch.cuda.current_stream(eng.device).wait_stream(side2)
            torch.cuda.synchronize(eng.device)
            for b in buckets:
                g_e2e[b] = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_e2e[b]):
                    nn_static(b)

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
            g_e2e[next(b for b in buckets if b >= n_req)].replay()
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
        n_man = len(sup.managers)
        for _ in range(min(args.presteps, 2000) * n_man):     # trees grow, searches desynchronise (untimed)
            pred = e2e_step(pred)
        # as many polls as the device loop ran advances (per manager), i.e. seconds of wall clock
        reps = max(200, args.steps * K // 2) * n_man
        m0 = sup.totals()
        h2d = d2h = 0
        t_poll = t_nn = 0.0
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            pred = e2e_step(pred)
        torch.cuda.synchronize()
        dt = all_max(time.perf_counter() - t0, world, eng.device)
        m1 = sup.totals()
        e2e_moves = all_sum(m1['moves'] - m0['moves'], world, eng.device)
        e2e_evals = all_sum(m1['evals'] - m0['evals'], world, eng.device)
        Gs = Gm * n_man
        e2e = {'value': e2e_moves / dt, 'unit': 'moves/s', 'h2d_bytes_per_step': int(h2d / reps) * n_man * K,
               'd2h_bytes_per_step': int(d2h / reps) * n_man * K, 'evals_per_s': e2e_evals / dt,
               'timed_region_s': dt, 'moves_in_window': e2e_moves,
               'mean_ply_start': m0['depth_sum'] / Gs, 'mean_ply_end': m1['depth_sum'] / Gs,
               'api': 'interface.Supervisor.poll (numpy in / numpy out, reference contract), network fed from pinned host memory; '
                      'the forward is a CUDA graph per batch size in steps of %d rows, the smallest that holds the rows handed out' % (Gm // 8),
               'managers': '%d x %d games (%s)' % (n_man, Gm, 'inline' if n_man == 1 else
                                                   'worker mode: engine step of one manager overlaps the network call of the other'),
               'polls': reps, 'rows_per_poll': Gm, 'bytes_are': 'per bench step = %d polls of each manager' % K,
               'host_ms_per_poll': {'poll': t_poll * 1e3 / reps, 'network_call': t_nn * 1e3 / reps}}
        del sup

    # ---- end of 'generation': gather the finished samples to rank 0 over NCCL (the only collective), timed on its own
    recs = np.concatenate([e.fetch_sample_records() for e in engines], axis=0)
    produced = int(recs.shape[0])
    if args.gather_records > 0:
        # a generation's worth of records (templates.py:171: 20 000 samples) split over the ranks: the records
        # this run produced, repeated up to this rank's share, so that the collective moves a real generation's bytes
        share = (args.gather_records + world - 1) // world
        if produced == 0:
            recs = np.zeros((share, eng.srec), dtype=np.int32)
        else:
            recs = np.concatenate([recs] * ((share + produced - 1) // produced), axis=0)[:share]
    recs_dev = torch.from_numpy(recs).to(eng.device)
    barrier(world)
    tg0 = time.perf_counter()
    gathered = gzdist.gather_sample_records(recs_dev, dst=0)
    torch.cuda.synchronize(eng.device)
    gather_s = all_max(time.perf_counter() - tg0, world, eng.device)
    n_gathered = int(gathered.shape[0]) if gathered is not None else 0

    # ---- the step after the gather (the trainer's DataCache.sync, nn/datacache.py:700-760): the gathered
    # generation's records -> planes + dense policy rows + value rows on the device (csrc/gz_training.cu)
    training = None
    if rank == 0 and gathered is not None and n_gathered > 0:
        try:
            from galvanise_zero_b200.datacache import TrainingBatcher
            tb = TrainingBatcher(eng.transformer, eng.c_transformer, eng.state_words, device=local)
            gathered = gathered.contiguous()
            for _ in range(3):
                bad0 = tb.arrays(gathered, plane_dtype=torch.bfloat16, check=True)['bad']
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(eng.device)
            reps_t = 10
            t0e.record()
            for _ in range(reps_t):
                arr = tb.arrays(gathered, plane_dtype=torch.bfloat16)
            t1e.record()
            torch.cuda.synchronize(eng.device)
            t_ms = t0e.elapsed_time(t1e) / reps_t
            out_bytes = arr['planes'][0].numel() * 2 + sum(p.shape[1] for p in arr['policies']) * 4 + arr['value'].shape[1] * 4
            alg = int(gathered.shape[1]) * 4 + out_bytes          # a record read once, its rows written once
            ach = alg * n_gathered / (t_ms * 1e-3) / 1e9
            # the same kernels over up to 2^20 output rows (every record repeated through the row index: the rows
            # written, 1.2 GB for reversi, no longer fit the L2): the figure the HBM roofline applies to
            big = None
            if n_gathered >= 1024:
                rows_big = max(1 << 14, min(1 << 20, int(4e9 // max(1, out_bytes))))      # at most ~4 GB of rows
                idx_big = (torch.arange(rows_big, device=eng.device, dtype=torch.int64) % n_gathered).to(torch.int32)
                for _ in range(2):
                    tb.arrays(gathered, plane_dtype=torch.bfloat16, index=idx_big)
                torch.cuda.synchronize(eng.device)
                t0e.record()
                for _ in range(5):
                    tb.arrays(gathered, plane_dtype=torch.bfloat16, index=idx_big)
                t1e.record()
                torch.cuda.synchronize(eng.device)
                b_ms = t0e.elapsed_time(t1e) / 5
                b_ach = alg * rows_big / (b_ms * 1e-3) / 1e9
                big = {'rows': rows_big, 'ms': b_ms, 'rows_per_s': rows_big / (b_ms * 1e-3), 'achieved': b_ach,
                       'unit': 'GB/s', 'frac': b_ach / load_peaks()['hbm_gbs']}
                del idx_big
            training = {'records': n_gathered, 'ms': t_ms, 'large_batch': big, 'records_per_s': n_gathered / (t_ms * 1e-3),
                        'bytes_per_record': alg, 'bound': 'hbm', 'achieved': ach, 'peak': load_peaks()['hbm_gbs'],
                        'unit': 'GB/s', 'frac': ach / load_peaks()['hbm_gbs'],
                        'kernels': 'rec_unpack_kernel + encode kernel (+ memset of the policy rows), bf16 planes',
                        'check_sample_failures': bad0}
        except Exception as exc:
            training = {'unavailable': repr(exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_reference_selfplay_mean(game, CPU_INLINE_SAMPLES)
        except Exception as exc:   # oracle/_ref missing on this box
            cpu = {'value': None, 'unit': 'moves/s', 'cores': os.cpu_count(), 'kind': 'reference',
                   'sample': 'unavailable: %s' % exc}

    if rank == 0:
        own_launches = advances * runner.kernels_per_step + (propnet_line['gpu_launches'] if propnet_line else 0)
        line = {
            'metric': 'PUCT self-play moves/sec', 'value': value, 'unit': 'moves/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'timed_region_s': ms_max * 1e-3,
            'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': preset['workload'], 'games_per_gpu': G, 'evals_per_move': PUCT_EVALS_PER_MOVE,
                       'network': 'random-init small tower, %.1f MFLOP/position' % (net.flops_per_position / 1e6),
                       'step': ('%d engine advances (select/backup + propnet + expand + encode) over %d games, each followed by '
                                'one network forward over the leaves that wait for it (compacted on the device, batch sizes '
                                'in steps of %d rows; see regime.network_rows_per_advance)' % (K, G, max(1, G // 32)))
                               if (P == 1 and not args.no_graph and not args.full_batches) else
                               ('%d engine advances (select/backup + propnet + expand + encode), each followed by one '
                                'network forward over %d leaves' % (K, G)),
                       'advances_per_step': K, 'ms_per_advance': ms_max / advances,
                       'regime': 'slots start at staggered plies (uniform in [0, %d) random plies), %d untimed '
                                 'advances, then the warm-up steps' % (preset['stagger_plies'], args.presteps)
                                 if not args.no_stagger else 'all slots start at the initial state',
                       'l2': 'tree arenas %.1f GB per GPU >> 126 MB L2' % (sum(e.arena_bytes for e in engines) / 1e9),
                       'presteps': args.presteps, 'cuda_graph': not args.no_graph,
                       'pipelines': '%d engine+network pipelines of %d games each on separate streams' % (P, G // P),
                       'parallelism': 'games sharded across %d GPU(s), no collective on the hot path' % world},
            'evals_per_s': evals / (ms_max * 1e-3), 'moves_in_window': moves, 'regime': regime,
            'roofline': roofline, 'network_roofline': network_roofline, 'kernel_times_ms': parts,
            'cpu_baseline': cpu, 'e2e': e2e,
            'gpu_launches': own_launches, 'clocks': clocks,
            'engine_totals': {k: t_after[k] for k in ('errors', 'halted', 'max_nodes', 'max_children', 'games', 'samples')},
            'training_batch': training,
            'sample_gather': {'records_on_rank0': n_gathered, 'record_bytes': int(recs.shape[1]) * 4 if recs.ndim == 2 else 0,
                              'bytes_on_rank0': n_gathered * (int(recs.shape[1]) * 4 if recs.ndim == 2 else 0),
                              'seconds': gather_s, 'produced_by_rank0_in_this_run': produced,
                              'padded_to_generation_size': args.gather_records if args.gather_records > 0 else None,
                              'collective': 'all_gather of counts + all_gather of padded records (NCCL over NVLink; gloo in the CPU tests)'},
            'propnet': propnet_line,
        }
        print(json.dumps(line), flush=True)


def run_reference_puct(args):
    """Reference arm: the reference's own C++ search on the host cores in the same regime as the GPU arm and the
    in-line cpu_baseline (staggered plies, warm-up, bounded window); one sample per 2 contract steps, each in a
    fresh process with its own seeds (the in-line baseline takes the first CPU_INLINE_SAMPLES of the same
    sequence), and the mean over the samples. With its 'propnet' half next to it."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    game = puct_game(args)
    t0 = time.perf_counter()
    n_samples = max(CPU_INLINE_SAMPLES, (max(1, args.steps) + 1) // 2)
    cpu = cpu_reference_selfplay_mean(game, n_samples)
    prop = None
    if not args.no_propnet and game == PUCT_GAME:
        prop = reference_propnet_line(args, steps=3)
    dt = time.perf_counter() - t0
    print(json.dumps({
        'impl': 'reference', 'metric': 'PUCT self-play moves/sec', 'value': cpu['value'], 'unit': 'moves/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3 / max(args.steps, 1), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': PUCT_PRESETS[game]['workload'], 'evals_per_move': PUCT_EVALS_PER_MOVE,
                   'note': 'host-core run of the reference C++ search; concurrent games bounded to what the '
                           'host threads hold (the GPU arm holds thousands per GPU)'},
        'evals_per_s': cpu['evals_per_s'], 'cpu_baseline': cpu,
        'e2e': {'value': cpu['value'], 'unit': 'moves/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'propnet': prop}),
        flush=True)


def reference_propnet_line(args, steps=None):
    """CPU reference arm of the propnet metric: the oracle port of the reference's state-machine loop on all host
    cores (ggplib is absent, so the reference state machine itself cannot be built)."""
    threads = os.cpu_count() or 1
    steps = steps or args.steps
    from galvanise_zero_b200 import tables
    from oracle import pnref
    blob = tables.blob_by_name(PROPNET_GAME)
    secs, tr, _ = pnref.playout_bench(blob, 64 * threads, 50, 0xB200, threads)
    rate = tr / max(secs, 1e-9)
    plies = 100
    slots = max(threads, int(rate * 3.0 / plies))       # ~3 s of CPU work per step
    for _ in range(min(args.warmup, 3)):
        pnref.playout_bench(blob, max(threads, slots // 8), plies, 0xB200, threads)
    t0 = time.perf_counter()
    total = 0
    for k in range(steps):
        secs, tr, _ = pnref.playout_bench(blob, slots, plies, 0xB200 + k, threads)
        total += tr
    dt = time.perf_counter() - t0
    val = total / dt
    cpu = {'value': val, 'unit': 'transitions/s', 'cores': threads, 'kind': 'port',
           'sample': '%s: %d slots x %d plies per step, %d steps, %d host threads (oracle/propnet_ref.c; ggplib '
                     'is absent so the reference state machine itself cannot be built)' % (
                         PROPNET_GAME, slots, plies, steps, threads)}
    return {
        'impl': 'reference', 'metric': 'propnet transitions/sec', 'value': val,
        'unit': 'transitions/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3 / max(steps, 1), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
        'config': {'workload': 'breakthrough 8x8 propnet, batched uniform-random playouts'},
        'cpu_baseline': cpu,
        'e2e': {'value': val, 'unit': 'transitions/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0}}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    print(json.dumps(reference_propnet_line(args)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=0, help='default: 20 (puct), 2000 (propnet)')
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='puct', choices=['propnet', 'puct'])
    ap.add_argument('--sub-iters', type=int, default=1)
    ap.add_argument('--presteps', type=int, default=1500, help='untimed engine advances before the warm-up steps')
    ap.add_argument('--advances-per-step', type=int, default=256,
                    help='puct workload: engine advances (+ network forwards) that make up one bench step')
    ap.add_argument('--propnet-launches-per-step', type=int, default=300,
                    help='puct workload: playout launches per step of the propnet half of the line')
    ap.add_argument('--no-propnet', action='store_true', help='puct workload: leave the propnet half out')
    ap.add_argument('--no-stagger', action='store_true', help='all slots start at the initial state')
    ap.add_argument('--gather-records', type=int, default=20000,
                    help='records in the timed end-of-generation gather (the run\'s own records repeated; 0: only those produced)')
    ap.add_argument('--group-min-batch', type=int, default=None,
                    help='propnet: batches of at least this many states use the group kernel (default 1024; -1: never)')
    ap.add_argument('--max-iters', type=int, default=0, help='engine step limit: controller iterations per game')
    ap.add_argument('--max-work', type=int, default=0, help='engine step limit: tree levels descended per game')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--pipelines', type=int, default=1)
    ap.add_argument('--staggered', action='store_true', help='(implied by --pipelines > 1)')
    ap.add_argument('--full-batches', action='store_true',
                    help='puct: the network runs over all G rows every advance instead of the compacted waiting rows')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-inline', action='store_true')
    ap.add_argument('--slots', type=int, default=0)
    ap.add_argument('--game', default='', help='another compiled game (propnet default: breakthrough, puct: reversi)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps <= 0:
        args.steps = 20 if args.workload == 'puct' else 2000
    if args.impl == 'reference':
        if args.workload == 'puct':
            run_reference_puct(args)
        else:
            run_reference(args)
        return
    rank, world, local = dist_setup(args.gpus)
    if args.workload == 'propnet':
        line = run_propnet(args, rank, world, local)
        if line is not None:
            print(json.dumps(line), flush=True)
    else:
        prop = None
        if not args.no_propnet and puct_game(args) == PUCT_GAME:
            # BASELINE's first metric, measured in the same run on the same GPUs
            prop = run_propnet(args, rank, world, local, steps=args.steps * args.propnet_launches_per_step)
        run_puct(args, rank, world, local, propnet_line=prop)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
