"""
Golden vectors for the wrapper-level boundary (SURVEY.md section 8 a9 / b), produced by the REFERENCE'S OWN
``DLWPTorchNN.predict_timeseries`` (DLWP/model/models_torch.py:325-379) imported on the shim, with its ``predict`` bound
to the oracle's ``unet2`` (the reference's network needs TensorFlow / a GPU build of these layers; the loop, the
step_sequence window and the output reshapes are the reference's own code).  Build container only:

    python tests/golden/make_golden_timeseries.py        # -> tests/golden/timeseries.npz (committed)

Case: C8 faces, 2 variables x time_dim 2 = 4 channels, base 8, batch 2, channels_first predictors (the layout the
reference's reshape assumes), float64 oracle network, weights ``make_unet2_params(4, 4, base=8, seed=5)``.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch._dynamo  # noqa: F401  (must be imported before the shim registers its `tensorflow` stand-in)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..', 'oracle'))
import tf_shim  # noqa: E402
import cs_oracle as O  # noqa: E402

tf_shim.load_reference_generators()                       # registers DLWP / DLWP.model stubs + DLWP.util on the shim
models_torch = importlib.import_module('DLWP.model.models_torch')

N, B, V, TD, BASE, SEED = 8, 2, 2, 2, 8, 5
C = V * TD
params = {k: v.double() for k, v in O.make_unet2_params(C, C, base=BASE, seed=SEED).items()}


def oracle_predict(p):
    """numpy (B, C, 6, N, N) -> numpy, through the oracle's channels_last unet2 (what `self.model(p)` computes)."""
    x = torch.from_numpy(np.asarray(p, dtype=np.float64)).permute(0, 2, 3, 4, 1)
    with torch.no_grad():
        y = O.unet2(params, x)
    return y.permute(0, 4, 1, 2, 3).numpy().astype(np.float32)


dlwp = models_torch.DLWPTorchNN(is_convolutional=True, is_recurrent=False, time_dim=TD, scaler_type=None,
                                scale_targets=False)
dlwp.predict = oracle_predict                             # everything else below runs the reference's own method

rng = np.random.default_rng(11)
predictors = rng.standard_normal((B, C, 6, N, N)).astype(np.float32)
out = {'predictors': predictors, 'meta': np.array([N, B, V, TD, BASE, SEED])}
for name, steps, seq, keep in (('plain5', 5, False, False), ('plain5_keep', 5, False, True), ('seq3', 3, True, False),
                               ('seq3_keep', 3, True, True), ('plain2', 2, False, False)):
    out[name] = dlwp.predict_timeseries(predictors.copy(), steps, step_sequence=seq, keep_time_dim=keep)
    print(name, out[name].shape, float(np.abs(out[name]).max()))
np.savez_compressed(os.path.join(HERE, 'timeseries.npz'), **out)
