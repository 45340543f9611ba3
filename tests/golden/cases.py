"""Case tables shared by make_golden.py (build container, needs the reference) and the tests (anywhere)."""

PAD_CASES = [(4, 1), (5, 2), (8, 3), (48, 1), (24, 1), (12, 1), (6, 2)]

# (name, ctor kwargs, batch, H (already padded input edge), Cin)
CONV_CASES = [
    ('k3_valid_flip', dict(filters=5, kernel_size=3), 2, 8, 4),
    ('k3_valid_noflip', dict(filters=5, kernel_size=3, flip_north_pole=False), 2, 8, 4),
    ('k3_indep_north', dict(filters=3, kernel_size=3, independent_north_pole=True, flip_north_pole=False), 1, 7, 2),
    ('k3_indep_north_flip', dict(filters=3, kernel_size=3, independent_north_pole=True, flip_north_pole=True), 1, 7, 2),
    ('k1_output', dict(filters=6, kernel_size=1), 2, 6, 8),
    ('k3_nobias', dict(filters=4, kernel_size=3, use_bias=False), 1, 6, 3),
    ('k3_same', dict(filters=4, kernel_size=3, padding='same'), 1, 6, 3),
    ('k3_same_stride2', dict(filters=4, kernel_size=3, padding='same', strides=2), 1, 7, 3),
    ('k3_valid_stride2', dict(filters=4, kernel_size=3, strides=2), 1, 8, 3),
    ('k3_dilation2', dict(filters=4, kernel_size=3, dilation_rate=2), 1, 9, 3),
    ('k23_rect', dict(filters=3, kernel_size=(2, 3)), 1, 7, 2),
    ('k5_valid', dict(filters=2, kernel_size=5), 1, 9, 2),
]


# device-side data feed (tests/golden/make_golden_feed.py): keyword arguments of oracle/cs_feed.generate per golden case
FEED_CASES = {'a': dict(input_slice=slice(0, 4), output_slice=slice(0, 3), t_in=2, t_out=2, interval=2),
              'b': dict(input_slice=slice(None), output_slice=slice(None), t_in=2, t_out=2, interval=1),
              'c': dict(input_slice=slice(1, 5, 2), output_slice=slice(3, 4), t_in=1, t_out=3, interval=3)}
FEED_SEQ_CASES = {'d': (dict(input_slice=slice(0, 4), output_slice=slice(0, 4), t_in=2, t_out=2, interval=1), 2),
                  'e': (dict(input_slice=slice(None), output_slice=slice(1, 3), t_in=1, t_out=2, interval=2), 3)}
