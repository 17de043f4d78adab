"""Import the UNMODIFIED reference from /root/reference (build container only).

Two compatibility patches, neither touching reference files:
 1. torch >= 1.5 ``_ConvNd.__init__`` needs ``padding_mode`` and tuple arguments, while
    ``model_zoo/nconv.py:23`` passes 10 positionals with int padding.
 2. ``model_zoo/nconv_sd.py:45,59,72`` and ``s2d_resnet.py:20`` call ``.cuda()``; on a CPU-only
    box that is made a no-op.
``/root/reference`` does not exist on the GPU box; there the byte-identical copy staged by ``oracle/make_ref.py`` under
``oracle/_ref`` (git-ignored, travels with the snapshot) is used.  With neither, everything that uses this module is skipped
(``have_reference()``).
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
REF = "/root/reference" if os.path.isdir("/root/reference/model_zoo") else _STAGED


def have_reference():
    return os.path.isdir(os.path.join(REF, "model_zoo"))


_done = False


def load():
    """Returns a namespace with the reference modules (nconv, nconv_sd, cdf, s2du, s2d, metrics, d2s)."""
    global _done
    import torch
    from torch.nn.modules import conv as _conv
    from torch.nn.modules.utils import _pair
    if not _done:
        orig = _conv._ConvNd.__init__

        def patched(self, in_channels, out_channels, kernel_size, stride, padding, dilation, transposed,
                    output_padding, groups, bias, padding_mode="zeros", device=None, dtype=None):
            return orig(self, in_channels, out_channels, _pair(kernel_size), _pair(stride), _pair(padding),
                        _pair(dilation), transposed, _pair(output_padding), groups, bias, padding_mode, device, dtype)

        _conv._ConvNd.__init__ = patched
        if not torch.cuda.is_available():
            torch.Tensor.cuda = lambda self, *a, **k: self
        if REF not in sys.path:
            sys.path.insert(0, REF)
        _done = True
    import model_zoo.nconv as nconv
    import model_zoo.nconv_sd as nconv_sd
    import model_zoo.confidence_depth_framework as cdf
    import model_zoo.s2d_u_resnet as s2du
    import model_zoo.s2d_resnet as s2d
    import metrics
    import dataloaders.dense_to_sparse as d2s
    return types.SimpleNamespace(nconv=nconv, nconv_sd=nconv_sd, cdf=cdf, s2du=s2du, s2d=s2d, metrics=metrics, d2s=d2s)


def load_trainer():
    """The reference's ``trainer`` module (trainer.py imports utils.py, which imports matplotlib: absent in this image and only used
    for the validation image grids, so it is stubbed in sys.modules).  Also returns the namespace of ``load()``."""
    ns = load()
    import importlib.machinery
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.figure", "matplotlib.backends", "matplotlib.backends.backend_agg"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = m
    plt = sys.modules["matplotlib.pyplot"]
    if not hasattr(plt, "cm"):
        plt.cm = types.SimpleNamespace(viridis=None, jet=None, binary=None)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib.backends.backend_agg"].FigureCanvasAgg = type("FigureCanvasAgg", (), {})
    sys.modules["matplotlib.figure"].Figure = type("Figure", (), {})
    sys.modules["matplotlib"].pyplot = plt
    import trainer
    return trainer, ns
