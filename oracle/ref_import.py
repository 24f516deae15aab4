"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference from /root/reference (this container only).

/root/reference does not exist on the GPU box: nothing executed by `-m gpu` tests, smoke() or bench.py may call
into this module.  It is used by (a) oracle/make_golden.py to generate the committed fixtures under
tests/golden/ and (b) `-m "not gpu"` tests that pin oracle/networks_oracle.py and oracle/fx_oracle.py against
the reference modules themselves when the reference tree is present.

Import recipe: shim dir first, then the reference package roots.
"""
import copy
import os
import sys

REFERENCE_ROOT = os.environ.get("MST_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mixing_style_transfer", "networks"))


def _prepend(path):
    if path in sys.path:
        sys.path.remove(path)
    sys.path.insert(0, path)


def import_reference_networks():
    """Returns the reference module `networks.architectures` (FXencoder, TCNModel, TCNBlock ...)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    # the product package also has a sub-package called `networks`; the reference one is a TOP-LEVEL package
    _prepend(os.path.join(REFERENCE_ROOT, "mixing_style_transfer"))
    _prepend(_SHIMS)
    import importlib

    mod = importlib.import_module("networks.architectures")
    if not os.path.abspath(mod.__file__).startswith(os.path.abspath(REFERENCE_ROOT)):
        raise RuntimeError(f"`networks` resolved to {mod.__file__}, not the reference")
    return mod


def import_reference_fx():
    """Returns the reference module `common_audioeffects` running on the pymixconsole/soxbindings stubs."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _prepend(os.path.join(REFERENCE_ROOT, "mixing_style_transfer", "mixing_manipulator"))
    _prepend(_SHIMS)
    import importlib

    return importlib.import_module("common_audioeffects")


def reference_configs():
    import yaml

    with open(os.path.join(REFERENCE_ROOT, "inference", "configs.yaml"), "r") as f:
        return yaml.full_load(f)


def build_reference_models(enc_sd=None, tcn_sd=None):
    """Construct the reference FXencoder / TCNModel exactly as inference/style_transfer.py:47-57 does."""
    arch = import_reference_networks()
    cfg = reference_configs()
    enc = arch.FXencoder(copy.deepcopy(cfg["Effects_Encoder"]["default"]))  # ctor mutates its config (:30)
    c = cfg["TCN"]["default"]
    tcn = arch.TCNModel(nparams=c["condition_dimension"], ninputs=2, noutputs=2, nblocks=c["nblocks"],
                        dilation_growth=c["dilation_growth"], kernel_size=c["kernel_size"],
                        channel_width=c["channel_width"], stack_size=c["stack_size"],
                        cond_dim=c["condition_dimension"], causal=c["causal"])
    if enc_sd is not None:
        enc.load_state_dict(enc_sd)
    if tcn_sd is not None:
        tcn.load_state_dict(tcn_sd)
    return enc.eval(), tcn.eval()


def import_reference_loader_utils():
    """The reference's data_loader/loader_utils.py (load_wav_segment ...), loaded straight from its file: the package
    `data_loader/__init__.py` pulls soundfile / librosa / pyloudnorm, none of which is in this image; loader_utils itself
    only needs `soundfile` for the writer helpers, so an empty stand-in module is registered for the import."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib.util
    import types

    if "soundfile" not in sys.modules:
        sys.modules["soundfile"] = types.ModuleType("soundfile")
    path = os.path.join(REFERENCE_ROOT, "mixing_style_transfer", "data_loader", "loader_utils.py")
    spec = importlib.util.spec_from_file_location("mst_reference_loader_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_reference_normalizer():
    """The reference's mixing_manipulator modules behind the input FX normaliser: (data_normalization, fx_utils,
    normalization_imager).  librosa / matplotlib / aubio / soundfile are empty stand-ins (oracle/shims/); `pyloudnorm` resolves
    to the RESTATED meter of oracle/norm_oracle.py, so what this pins is the reference's own arithmetic around it
    (lufs_normalize, normalize_imager, Audio_Effects_Normalizer.normalize_audio_per_effect)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import importlib
    import types

    _prepend(os.path.join(REFERENCE_ROOT, "mixing_style_transfer", "mixing_manipulator"))
    _prepend(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # `oracle` package for the pyloudnorm shim
    _prepend(_SHIMS)
    if "soundfile" not in sys.modules:
        sys.modules["soundfile"] = types.ModuleType("soundfile")
    return (importlib.import_module("data_normalization"), importlib.import_module("fx_utils"),
            importlib.import_module("normalization_imager"))
