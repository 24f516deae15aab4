"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for `pytorch_lightning`.

The reference's TCNModel derives from `pl.LightningModule` only to get `save_hyperparameters()`
(mixing_style_transfer/networks/architectures.py:75-76,111). The package is absent from this image, so the
oracle harness places this shim on sys.path *before* importing the reference. Never imported by the product.
"""
import inspect

import torch


class _HParams(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)  # KeyError here would break copy.deepcopy


class LightningModule(torch.nn.Module):
    def save_hyperparameters(self):
        frame = inspect.currentframe().f_back  # caller = TCNModel.__init__
        args = inspect.getargvalues(frame)
        self.hparams = _HParams({k: args.locals[k] for k in args.args if k != "self"})
