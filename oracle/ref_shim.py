"""TEST INFRASTRUCTURE ONLY -- not part of the product path.

Imports the UNMODIFIED reference module ``/root/reference/spock_reg_model.py``
in this (CPU-only) container so that golden vectors can be generated from the
reference itself.  ``/root/reference`` does not exist on the GPU box, so nothing
under ``tests/ -m gpu``, ``smoke()`` or ``bench.py`` may import this file; only
``oracle/make_golden.py`` and the CPU-side oracle-vs-reference tests do.

Three stubs are installed in ``sys.modules`` because the reference's imports
(spock_reg_model.py:6-20) name packages that are absent from this image:
  * ``pytorch_lightning`` (LightningModule / Trainer / seed_everything /
    utilities.parsing.AttributeDict, the latter is a global inside the pickles),
  * ``torch._six``       (removed from torch >= 2.0; only ``inf`` is used),
  * ``matplotlib``       (``mpl.use('agg')`` and ``pyplot`` are import-time only).
"""
import io
import os
import pickle
import pickletools
import sys
import types
import zipfile

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("BNN_REFERENCE_ROOT", "/root/reference")

# The only globals a SWAG checkpoint may name (audited by audit_pickle()).
ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"),
    ("torch._utils", "_rebuild_tensor_v2"),
    ("torch", "FloatStorage"),
    ("pytorch_lightning.utilities.parsing", "AttributeDict"),
}


class AttributeDict(dict):
    """Stand-in for pytorch_lightning.utilities.parsing.AttributeDict."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, val):
        self[key] = val


class _LightningModule(nn.Module):
    """Minimal LightningModule: an nn.Module with the attributes the reference reads."""

    def __init__(self):
        super().__init__()
        self.current_epoch = 0
        self.global_step = 0

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def save_hyperparameters(self, *a, **k):
        return None


def _seed_everything(seed):
    import random

    import numpy as np

    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    return seed


def install_stubs():
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _LightningModule
        pl.seed_everything = _seed_everything
        pl.Trainer = type("Trainer", (), {})
        util = types.ModuleType("pytorch_lightning.utilities")
        parsing = types.ModuleType("pytorch_lightning.utilities.parsing")
        parsing.AttributeDict = AttributeDict
        util.parsing = parsing
        pl.utilities = util
        sys.modules["pytorch_lightning"] = pl
        sys.modules["pytorch_lightning.utilities"] = util
        sys.modules["pytorch_lightning.utilities.parsing"] = parsing
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six")
        six.inf = float("inf")
        sys.modules["torch._six"] = six
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def import_reference():
    """Return the reference's ``spock_reg_model`` module, imported unmodified."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise FileNotFoundError(f"reference tree not present at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import spock_reg_model  # the reference's, by path

    assert os.path.dirname(spock_reg_model.__file__) == REFERENCE_ROOT
    return spock_reg_model


def audit_pickle(path):
    """Static opcode scan of a torch zip checkpoint: return the set of globals it names."""
    found = set()
    with zipfile.ZipFile(path) as zf:
        names = [n for n in zf.namelist() if n.endswith("data.pkl")]
        assert len(names) == 1, names
        data = zf.read(names[0])
    strings = []
    for op, arg, _pos in pickletools.genops(data):
        if op.name == "GLOBAL":
            mod, name = arg.split(" ")
            found.add((mod, name))
        elif op.name in ("SHORT_BINUNICODE", "BINUNICODE", "UNICODE"):
            strings.append(arg)
        elif op.name == "STACK_GLOBAL":
            found.add((strings[-2], strings[-1]))
    return found


def load_checkpoint_dict(path):
    """Load a reference SWAG pickle after the opcode audit passes."""
    globs = audit_pickle(path)
    bad = globs - ALLOWED_GLOBALS
    if bad:
        raise pickle.UnpicklingError(f"unexpected globals in {path}: {sorted(bad)}")
    install_stubs()
    return torch.load(path, map_location="cpu", weights_only=False)


def load_reference_swag(path):
    """reference load_swag() (spock_reg_model.py:922-967) with the audited loader."""
    ref = import_reference()
    globs = audit_pickle(path)
    bad = globs - ALLOWED_GLOBALS
    if bad:
        raise pickle.UnpicklingError(f"unexpected globals in {path}: {sorted(bad)}")
    _orig = torch.load
    try:
        torch.load = lambda p, *a, **k: _orig(p, map_location="cpu", weights_only=False)
        model = ref.load_swag(path)
    finally:
        torch.load = _orig
    return model


def pretrained_path(seed):
    return os.path.join(
        REFERENCE_ROOT,
        "pretrained",
        "steps=300000_megno=0_angles=1_power=0_hidden=40_latent=20_nommr=1_nonan=1_noeplusminus=1_v50_"
        f"{seed}_output.pkl",
    )
