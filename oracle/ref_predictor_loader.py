"""TEST INFRASTRUCTURE ONLY — imports the REAL reference `nnUNetPredictor`
(`/root/reference/dinounet/inference/predict_from_raw_data.py`) on CPU, in the build container only.

The module pulls in most of the reference's data stack; the third-party packages that are absent here
(acvl_utils, batchgenerators, SimpleITK, nibabel, skimage, …) are replaced by inert stub modules through a meta-path
finder.  None of them is executed by the sliding-window methods except `pad_nd_image` (acvl-utils), for which the
restatement in `oracle/sliding_window_oracle.py` is injected.
"""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types
import typing

from . import ref_loader
from . import sliding_window_oracle as SWO

_STUB_ROOTS = {"acvl_utils", "batchgenerators", "batchgeneratorsv2", "SimpleITK", "nibabel", "skimage", "tifffile",
               "imagecodecs", "dicom2nifti", "seaborn", "matplotlib", "graphviz", "blosc2", "medpy"}


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {"__init__": lambda self, *a, **k: None})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def load_reference_predictor_class():
    """Returns the unmodified `nnUNetPredictor` class of the reference."""
    name = "dinounet.inference.predict_from_raw_data"
    if name in sys.modules:
        return sys.modules[name].nnUNetPredictor
    ref_loader._install_shims()
    sys.meta_path.append(_Finder())
    ff = _StubModule("batchgenerators.utilities.file_and_folder_operations")
    ff.__path__ = []
    ff.__dict__.update(dict(List=typing.List, Tuple=typing.Tuple, Union=typing.Union, Optional=typing.Optional, os=os,
                            join=os.path.join, isfile=os.path.isfile, isdir=os.path.isdir,
                            maybe_mkdir_p=lambda p: os.makedirs(p, exist_ok=True)))
    sys.modules[ff.__name__] = ff
    pad = _StubModule("acvl_utils.cropping_and_padding.padding")
    pad.__path__ = []

    def pad_nd_image(image, new_shape=None, mode="constant", kwargs=None, return_slicer=False,
                     shape_must_be_divisible_by=None):
        assert mode == "constant" and return_slicer and shape_must_be_divisible_by is None
        return SWO.pad_nd_image(image, new_shape)

    pad.pad_nd_image = pad_nd_image
    sys.modules[pad.__name__] = pad
    helper = sys.modules["dynamic_network_architectures.building_blocks.helper"]
    for n in ("convert_dim_to_conv_op", "get_matching_instancenorm", "get_matching_batchnorm", "get_matching_dropout",
              "maybe_convert_scalar_to_list"):
        if not hasattr(helper, n):
            setattr(helper, n, lambda *a, **k: None)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        mod = importlib.import_module(name)
    return mod.nnUNetPredictor
