"""Fixture configurations shared by oracle/gen_golden.py and the tests (test infrastructure)."""

TINY_ADM = dict(
    image_size=16,
    num_channels=32,
    channel_mult=(1, 2),
    num_res_blocks=1,
    attention_resolutions=(8,),
    num_head_channels=16,
    resblock_updown=True,
    use_scale_shift_norm=True,
)

# The imagenet_256x256 card's structure (reference azula/plugins/adm/cards.yaml:36-50) at a
# CPU-friendly width: same depth / multipliers / attention placement / head width, model
# channels 64, 64x64 input.
MID_ADM = dict(
    image_size=64,
    num_channels=64,
    channel_mult=(1, 1, 2, 2, 4, 4),
    num_res_blocks=2,
    attention_resolutions=(8, 4, 2),
    num_head_channels=64,
    resblock_updown=True,
    use_scale_shift_norm=True,
)

# The card itself (cards.yaml:36-50): 552.8 M parameters.
IMAGENET_256 = dict(
    discrete_schedule="linear",
    discrete_steps=1000,
    attention_resolutions=(32, 16, 8),
    channel_mult=(1, 1, 2, 2, 4, 4),
    image_size=256,
    num_channels=256,
    num_classes=None,
    num_head_channels=64,
    num_res_blocks=2,
    resblock_updown=True,
    use_scale_shift_norm=True,
)
