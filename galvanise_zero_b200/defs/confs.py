"""Configuration surface of the hot path, accepted verbatim from reference users.

Field names, defaults and meaning follow ``src/ggpzero/defs/confs.py:9-123`` of the reference
(``PUCTEvaluatorConfig``, ``PUCTPlayerConfig``, ``SelfPlayConfig``); the engines receive them as
``attr.asdict(conf)`` dicts exactly as ``src/ggpzero/util/cppinterface.py:167-169,209`` does.
Self-play template values follow ``src/ggpzero/defs/templates.py:73-129``.
"""
import attr


def _a(default):
    return attr.ib(default=default)


@attr.s
class PUCTEvaluatorConfig(object):
    verbose = _a(False)
    puct_constant = _a(0.85)
    puct_constant_root = _a(2.5)
    dirichlet_noise_pct = _a(0.25)              # < 0: off
    noise_policy_squash_pct = _a(-1)
    noise_policy_squash_prob = _a(0.05)
    choose = _a("choose_top_visits")            # | "choose_temperature"
    max_dump_depth = _a(2)
    random_scale = _a(0.5)
    temperature = _a(1.0)
    depth_temperature_start = _a(5)
    depth_temperature_increment = _a(0.5)
    depth_temperature_stop = _a(10)
    depth_temperature_max = _a(5.0)
    fpu_prior_discount = _a(0.25)
    fpu_prior_discount_root = _a(0.25)
    think_time = _a(10.0)
    converged_visits = _a(5000)
    top_visits_best_guess_converge_ratio = _a(0.8)
    evaluation_multiplier_to_convergence = _a(1.0)
    batch_size = _a(32)
    use_legals_count_draw = _a(-1)
    backup_finalised = _a(False)
    lookup_transpositions = _a(False)


@attr.s
class PUCTPlayerConfig(object):
    name = _a("Player")
    verbose = _a(False)
    playouts_per_iteration = _a(800)
    playouts_per_iteration_noop = _a(1)
    generation = _a("latest")
    evaluator_config = attr.ib(default=attr.Factory(PUCTEvaluatorConfig))


@attr.s
class SelfPlayConfig(object):
    oscillate_sampling_pct = _a(0.25)
    temperature_for_policy = _a(1.0)
    puct_config = attr.ib(default=attr.Factory(PUCTEvaluatorConfig))
    evals_per_move = _a(800)
    resign0_score_probability = _a(0.9)
    resign0_pct = _a(0.5)
    resign1_score_probability = _a(0.975)
    resign1_pct = _a(0.1)
    run_to_end_pct = _a(0.2)
    run_to_end_evals = _a(42)
    run_to_end_puct_config = attr.ib(default=attr.Factory(PUCTEvaluatorConfig))
    run_to_end_early_score = _a(0.01)
    run_to_end_minimum_game_depth = _a(30)
    abort_max_length = _a(-1)


# ---- templates (src/ggpzero/defs/templates.py:73-129)
def base_puct_config(**kwds):
    conf = PUCTEvaluatorConfig(
        verbose=False, puct_constant=0.85, puct_constant_root=0.85, dirichlet_noise_pct=-1,
        fpu_prior_discount=0.25, fpu_prior_discount_root=0.25, choose="choose_temperature",
        temperature=1.0, depth_temperature_max=5.0, depth_temperature_start=2,
        depth_temperature_increment=0.2, depth_temperature_stop=6, random_scale=0.95,
        max_dump_depth=0, top_visits_best_guess_converge_ratio=0.85, think_time=-1,
        converged_visits=1, evaluation_multiplier_to_convergence=2.0, batch_size=1)
    for k, v in kwds.items():
        if not hasattr(conf, k):
            raise AttributeError('PUCTEvaluatorConfig has no field %r' % k)
        setattr(conf, k, v)
    return conf


def selfplay_config_template(evals_per_move=100):
    return SelfPlayConfig(
        oscillate_sampling_pct=0.25, temperature_for_policy=1.0,
        puct_config=base_puct_config(dirichlet_noise_pct=0.25), evals_per_move=evals_per_move,
        resign0_score_probability=0.1, resign0_pct=0.99,
        resign1_score_probability=0.025, resign1_pct=0.95,
        run_to_end_pct=0.01, run_to_end_evals=32,
        run_to_end_puct_config=base_puct_config(dirichlet_noise_pct=0.15, random_scale=0.75),
        run_to_end_early_score=0.01, run_to_end_minimum_game_depth=30, abort_max_length=-1)
