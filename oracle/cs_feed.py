"""
CPU oracle (TEST INFRASTRUCTURE ONLY) for the device-side data feed: a numpy restatement of
``ArrayDataGenerator.generate`` (DLWP/model/generators.py:872-984) for the configuration every cubed-sphere training script
uses -- convolutional model, no time axis kept, no sequence, channels_last, insolation and constants supplied
(Azure/train_cs.py:140-175) -- pinned bit-for-bit against the reference's own class executed on the shim
(tests/golden/feed.npz, made by tests/golden/make_golden_feed.py).

    array            (time, varlev, 6, N, N)   predictor data
    insolation_array (time, 6, N, N)
    constants        (n_const, 6, N, N)

generate() -> ([p, constants_batch], targets) with (generators.py:880-899, 943-984)
    p[b, f, i, j, t*(V_in+1) + v]  = array[s_b + t*interval, input_slice[v], f, i, j]            v <  V_in
                                   = insolation_array[s_b + t*interval, f, i, j]                 v == V_in
    targets[b, f, i, j, t*V_out + v] = array[s_b + interval*(T_in + t), output_slice[v], f, i, j]
The Keras model concatenates [p, constants] on the channel axis (train_cs.py:396-407); ``generate`` here returns that
concatenation as one tensor, which is what the engine's model takes.
"""
import numpy as np


def generate_sequence(array, samples, input_slice, output_slice, t_in, t_out, interval, sequence, insolation_array,
                      constants=None):
    """``sequence=S`` mode (generators.py:884-893, 902-930, 964-983): the targets are S consecutive forecast steps and the
    inputs are [p, insolation of steps 1..S-1, constants].  Returns (x, solars, targets): x as in ``generate`` (predictors of
    step 0 with the constants appended), solars[s-1] = (B, T_in, 6, N, N, 1) insolation of the input times of step s,
    targets[s] = (B, 6, N, N, T_out*V_out) with channel t*V_out+v = array[s_b + interval*(T_in + T_out*s + t), v]."""
    samples = np.asarray(samples, dtype=np.int64)
    x, _ = generate(array, samples, input_slice, output_slice, t_in, t_out, interval, insolation_array, constants)
    solars = []
    for s in range(1, sequence):
        sol = np.concatenate([insolation_array[samples + interval * (t_in * s + k), np.newaxis, np.newaxis]
                              for k in range(t_in)], axis=1)                               # (B, T_in, 1, 6, N, N)
        solars.append(np.ascontiguousarray(sol.transpose((0, 1) + tuple(range(3, sol.ndim)) + (2,))))
    targets = [generate(array, samples + interval * t_out * s, input_slice, output_slice, t_in, t_out, interval)[1]
               for s in range(sequence)]
    return x, solars, targets


def generate(array, samples, input_slice, output_slice, t_in, t_out, interval, insolation_array=None, constants=None):
    samples = np.asarray(samples, dtype=np.int64)
    n = len(samples)
    p = np.concatenate([array[samples + k * interval][:, input_slice][:, np.newaxis] for k in range(t_in)], axis=1)
    if insolation_array is not None:
        insol = np.concatenate([insolation_array[samples + k * interval, np.newaxis, np.newaxis] for k in range(t_in)], axis=1)
        p = np.concatenate([p, insol], axis=2)
    p = p.reshape((n, -1) + array.shape[2:])                       # (B, T*(V+1), 6, N, N)
    t = np.concatenate([array[samples + interval * (t_in + k)][:, np.newaxis][:, :, output_slice] for k in range(t_out)], axis=1)
    t = t.reshape((n, -1) + array.shape[2:])
    perm = (0,) + tuple(range(2, p.ndim)) + (1,)
    p, t = p.transpose(perm), t.transpose(perm)                    # channels_last
    if constants is not None:
        c = np.repeat(np.expand_dims(constants, axis=0), n, axis=0).transpose(perm)
        p = np.concatenate([p, c], axis=-1)
    return np.ascontiguousarray(p), np.ascontiguousarray(t)
