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
