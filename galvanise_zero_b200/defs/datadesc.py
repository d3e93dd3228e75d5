"""Sample / generation descriptors — same fields as ``src/ggpzero/defs/datadesc.py:8-36,56-94``."""
import attr


@attr.s
class Sample(object):
    state = attr.ib(default=attr.Factory(list))          # tuple/list of 0/1, base-index order
    prev_states = attr.ib(default=attr.Factory(list))    # first element = immediate parent
    policies = attr.ib(default=attr.Factory(list))       # per role: [(legal_index, prob), ...]
    final_score = attr.ib(default=attr.Factory(list))
    depth = attr.ib(default=0)
    game_length = attr.ib(default=0)
    match_identifier = attr.ib(default="")
    has_resigned = attr.ib(default=False)
    resign_false_positive = attr.ib(default=False)
    starting_sample_depth = attr.ib(default=0)
    resultant_puct_score = attr.ib(default=attr.Factory(list))
    resultant_puct_visits = attr.ib(default=800)


@attr.s
class GenerationDescription(object):
    game = attr.ib(default="breakthrough")
    name = attr.ib(default="v6_123")
    date_created = attr.ib(default='2018-01-24 22:28')
    channel_last = attr.ib(default=False)
    multiple_policy_heads = attr.ib(default=False)
    num_previous_states = attr.ib(default=0)
    transformer_description = attr.ib(default=None)
    draw_head = attr.ib(default=False)
    trained_losses = attr.ib(default='not set')
    trained_validation_losses = attr.ib(default='not set')
    trained_policy_accuracy = attr.ib(default='not set')
    trained_value_accuracy = attr.ib(default='not set')


def generation_for(game, num_previous_states=1, draw_head=None):
    """The shape every shipped generation of the reference uses (data/*/generations/*.json):
    channels-first, one policy head per role, one previous state; draw head for reversi/chess."""
    if draw_head is None:
        draw_head = game in ('reversi', 'reversi_10x10', 'chess_15d', 'draughts_killer_10x10')
    return GenerationDescription(game=game, name='b200_synthetic', channel_last=False,
                                 multiple_policy_heads=True,
                                 num_previous_states=num_previous_states, draw_head=draw_head)


@attr.s
class GenerationSamples(object):
    """``src/ggpzero/defs/datadesc.py:39-53`` — one generation's samples as stored on disk / sent on the wire."""
    game = attr.ib(default="game")
    date_created = attr.ib(default='2018-01-24 22:28')
    with_generation = attr.ib(default="v6_123")
    num_samples = attr.ib(default=1024)
    samples = attr.ib(default=attr.Factory(list))
