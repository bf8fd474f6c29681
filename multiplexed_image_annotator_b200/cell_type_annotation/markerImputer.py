"""`MarkerImputer`: the reference's MAE marker imputation (cta/markerImputer.py:258-329) behind the
same constructor and `impute(data, batch_size)` call, running ribca_mae_impute on the device.

Checkpoints are read from the same CWD-relative paths as the reference (`*_impute.pth`,
markerImputer.py:260-271).  The constant 0.1 / 0.8 "noise" of the reference makes its random masking
a fixed compaction (kept tokens = present channels), so no sort or RNG runs on the device.
"""
import os

import torch

from ..engine import MaeEngine
from ..weights import MAE_SPECS, MODEL_DIR, load_checkpoint

_STATE_OVERRIDES = {}      # panel -> state dict, used by tests / benchmarks instead of a file


def register_state(panel: str, state_dict) -> None:
    """Provide imputer weights in memory (random-init benchmarks have no checkpoint files)."""
    _STATE_OVERRIDES[panel] = state_dict


class MarkerImputer():
    def __init__(self, channel_index, device, panel=""):
        if panel not in MAE_SPECS:
            raise ValueError("Panel not found")
        spec = MAE_SPECS[panel]
        path = os.path.join(MODEL_DIR, spec.ckpt)
        if panel in _STATE_OVERRIDES:
            state = _STATE_OVERRIDES[panel]
        elif os.path.exists(path):
            state = load_checkpoint(path)
        else:
            raise ValueError("Panel not found")          # markerImputer.py:275-276
        self.device = device
        self.shape = spec.grid
        self.channel_index = list(channel_index)
        self.channel_number = spec.channels
        self.model = MaeEngine(spec, state, device=device)

    def impute(self, data, batch_size=1, precision=None):
        """data: (N, C_panel, 40, 40) float32.  A CUDA tensor is imputed in place; a CPU tensor makes
        the round trip the reference makes (markerImputer.py:301,317).  `batch_size` is accepted for
        signature compatibility; the engine picks its own chunking.  `precision` (not in the reference) selects a
        higher-precision pass for the exact-label re-evaluation."""
        if precision is not None and precision not in self.model.PRECISIONS:
            precision = "bf16x3"
        if data.is_cuda:
            return self.model.impute(data, self.channel_index, precision=precision)
        dev = data.to(self.device, torch.float32).contiguous()
        self.model.impute(dev, self.channel_index, precision=precision)
        data.copy_(dev.cpu())
        return data
