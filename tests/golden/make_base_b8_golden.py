"""Writes tests/golden/base_b8_predict_<steps>.npz: the ORACLE's output for the configuration
bench.py measures (BASELINE config 3: base_with_context, batch of 8 segments, CFG weight 2.0; fp32
graph as written) on the synthetic weights (seed 0) and the mixed batch of
tests/helpers.base_b8_batch, with the library's Philox noise (oracle/philox.py, seed 17) so the
CUDA path can be run from the seed alone.  ~4 minutes of 8-core CPU for 20 steps.
Usage: python tests/golden/make_base_b8_golden.py [steps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from music_spectrogram_diffusion_b200 import config, weights  # noqa: E402
from oracle import msd_oracle as O, philox  # noqa: E402
from tests import helpers as H  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
SEED, BATCH_SEED = 17, 321
torch.set_num_threads(len(os.sched_getaffinity(0)))
t5 = config.t5_base()
lengths = dict(config.TASK_FEATURE_LENGTHS_CONTEXT)
params = O.params_to(weights.synthetic_params(t5, 2048, 256, 256, seed=0))
b = H.base_b8_batch(lengths, BATCH_SEED)
oc = O.OracleConfig(num_steps=steps, eval_condition_weight=2.0)
shape = (8, 256, 128)
z0 = torch.from_numpy(philox.init_z(SEED, shape))
noise = lambda i: torch.from_numpy(philox.step_noise(SEED, i, shape))
batch = {k: torch.from_numpy(b[k]) for k in ('encoder_input_tokens', 'encoder_continuous_inputs',
                                             'encoder_continuous_mask')}
t0 = time.time()
with torch.no_grad():
  mel, _ = O.predict_batch_with_aux(params, oc, batch, z0, noise)
print(f'{steps} steps in {time.time() - t0:.0f} s; mel mean {float(mel.mean()):.4f}')
np.savez_compressed(
    os.path.join(os.path.dirname(os.path.abspath(__file__)), f'base_b8_predict_{steps}.npz'),
    mel=mel.numpy(), seed=SEED, steps=steps, cond_weight=2.0, batch_seed=BATCH_SEED, weight_seed=0)
