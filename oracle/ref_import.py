"""Import the UNMODIFIED Slot-VPS reference in place (test infrastructure only).

TEST INFRASTRUCTURE -- never imported by the product package ``slotvps_b200``.

The reference (``/root/reference``, read-only, Samsung non-commercial licence: nothing of it is
copied into this repo) is a fork of mmdetection that needs ``mmcv``, ``timm``, ``panopticapi``,
``pycocotools``, ``terminaltables``, ``imagecorruptions`` and three compiled CUDA extensions,
none of which are installed here.  None of those contribute arithmetic to the hot path
(SURVEY.md section 8c), so this module registers tiny stand-ins in ``sys.modules`` and then imports
``mmdet.models`` from the reference tree itself.  It is used by

* ``tests/golden/make_golden.py``  -- to generate the committed golden vectors, and
* ``tests/test_oracle_vs_reference.py`` -- live oracle-vs-reference checks, skipped when
  ``/root/reference`` is absent (the GPU box).

Nothing under ``-m gpu``, ``bench.py`` or ``__graft_entry__.smoke()`` may call this module.
"""
import os
import runpy
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("SLOTVPS_REFERENCE_ROOT", "/root/reference")
CONFIG = "configs/cityscapes/r50_fpn_slotvps.py"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mmdet"))


class AttrDict(dict):
    """dict with attribute access (stands in for mmcv.ConfigDict / easydict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def to_attr(obj):
    if isinstance(obj, dict):
        return AttrDict({k: to_attr(v) for k, v in obj.items()})
    if isinstance(obj, list):
        return [to_attr(v) for v in obj]
    return obj


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "_slotvps_stub", False):
        return

    # ---- mmcv ---------------------------------------------------------------------------
    def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
        if distribution == "uniform":
            nn.init.kaiming_uniform_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        else:
            nn.init.kaiming_normal_(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def xavier_init(module, gain=1, bias=0, distribution="normal"):
        if distribution == "uniform":
            nn.init.xavier_uniform_(module.weight, gain=gain)
        else:
            nn.init.xavier_normal_(module.weight, gain=gain)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def constant_init(module, val, bias=0):
        if getattr(module, "weight", None) is not None:
            nn.init.constant_(module.weight, val)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def normal_init(module, mean=0, std=1, bias=0):
        nn.init.normal_(module.weight, mean, std)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def uniform_init(module, a=0, b=1, bias=0):
        nn.init.uniform_(module.weight, a, b)
        if getattr(module, "bias", None) is not None:
            nn.init.constant_(module.bias, bias)

    def is_str(x):
        return isinstance(x, str)

    def is_list_of(seq, expected_type):
        return isinstance(seq, list) and all(isinstance(i, expected_type) for i in seq)

    class _Hook:
        pass

    class DataContainer:
        def __init__(self, data, *a, **k):
            self.data = data

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            raise RuntimeError("stubbed dependency called")

    def _noop(*a, **k):
        return None

    mmcv = _mod("mmcv", is_str=is_str, is_list_of=is_list_of, _slotvps_stub=True,
                imread=_noop, imwrite=_noop, imresize=_noop, imrescale=_noop, imflip=_noop,
                impad=_noop, impad_to_multiple=_noop, imnormalize=_noop, bgr2hsv=_noop,
                hsv2bgr=_noop, ProgressBar=_Anything, Config=_Anything, mkdir_or_exist=_noop,
                dump=_noop, load=_noop, track_iter_progress=_noop, imshow_det_bboxes=_noop,
                concat_list=_noop, slice_list=_noop, is_tuple_of=lambda s, t: isinstance(s, tuple))
    mmcv.cnn = _mod("mmcv.cnn", kaiming_init=kaiming_init, xavier_init=xavier_init,
                    constant_init=constant_init, normal_init=normal_init, uniform_init=uniform_init)
    mmcv.cnn.weight_init = _mod("mmcv.cnn.weight_init", kaiming_init=kaiming_init,
                                xavier_init=xavier_init, constant_init=constant_init,
                                normal_init=normal_init, uniform_init=uniform_init)
    get_dist_info = lambda: (0, 1)  # noqa: E731
    mmcv.runner = _mod("mmcv.runner", OptimizerHook=_Hook, Hook=_Hook, load_checkpoint=_noop,
                       load_state_dict=_noop, get_dist_info=get_dist_info, Runner=_Anything,
                       DistSamplerSeedHook=_Hook, obj_from_dict=_noop)
    mmcv.runner.utils = _mod("mmcv.runner.utils", get_dist_info=get_dist_info)
    mmcv.parallel = _mod("mmcv.parallel", DataContainer=DataContainer, collate=_noop, scatter=_noop,
                         MMDataParallel=_Anything, MMDistributedDataParallel=_Anything)
    mmcv.parallel.data_container = _mod("mmcv.parallel.data_container", DataContainer=DataContainer)
    mmcv.utils = _mod("mmcv.utils", is_str=is_str)

    # ---- timm (only DropPath is constructed, and only when drop_path > 0) ---------------
    class DropPath(nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    def to_2tuple(x):
        return x if isinstance(x, tuple) else (x, x)

    def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return nn.init.trunc_normal_(t, mean, std, a, b)

    timm = _mod("timm")
    timm.models = _mod("timm.models")
    timm.models.layers = _mod("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple,
                              trunc_normal_=trunc_normal_)

    # ---- panopticapi: the standard base-256 id<->rgb packings ----------------------------
    def rgb2id(color):
        if isinstance(color, np.ndarray) and color.ndim == 3:
            c = color.astype(np.int32) if color.dtype == np.uint8 else color
            return c[:, :, 0] + 256 * c[:, :, 1] + 256 * 256 * c[:, :, 2]
        return int(color[0] + 256 * color[1] + 256 * 256 * color[2])

    def id2rgb(id_map):
        if isinstance(id_map, np.ndarray):
            m = id_map.copy()
            rgb = np.zeros(tuple(list(id_map.shape) + [3]), dtype=np.uint8)
            for i in range(3):
                rgb[..., i] = m % 256
                m //= 256
            return rgb
        out = []
        for _ in range(3):
            out.append(id_map % 256)
            id_map //= 256
        return out

    pan = _mod("panopticapi")
    pan.utils = _mod("panopticapi.utils", rgb2id=rgb2id, id2rgb=id2rgb, IdGenerator=_Anything)

    # ---- misc never-called deps ------------------------------------------------------------
    pc = _mod("pycocotools")
    pc.coco = _mod("pycocotools.coco", COCO=_Anything)
    pc.cocoeval = _mod("pycocotools.cocoeval", COCOeval=_Anything)
    pc.mask = _mod("pycocotools.mask")
    _mod("terminaltables", AsciiTable=_Anything)
    _mod("imagecorruptions", corrupt=_noop)
    # compiled extensions (DCN / focal loss): present as empty modules, never executed
    for ext in ("mmdet.ops.dcn.deform_conv_cuda", "mmdet.ops.dcn.deform_pool_cuda",
                "mmdet.ops.sigmoid_focal_loss.sigmoid_focal_loss_cuda"):
        _mod(ext)


_CACHE = {}


def import_reference():
    """Return the reference's ``mmdet.models`` package (imported from REFERENCE_ROOT)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if "models" not in _CACHE:
        _install_stubs()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        import mmdet.models as models  # noqa: E402  (the reference's package)
        _CACHE["models"] = models
    return _CACHE["models"]


def import_reference_tools():
    """The reference's evaluation helper ``tools.dataset.cityscapes_vps.CityscapesVps`` (for get_unified_pan_result,
    tools/dataset/cityscapes_vps.py:214) with its dataset constants set as the shipped test yaml does
    (num_seg_classes=19, num_classes=9).  Stubs (names only): easydict, cv2, matplotlib; ``collections.Sequence``."""
    import_reference()
    if "tools_cls" not in _CACHE:
        import collections
        import collections.abc
        if not hasattr(collections, "Sequence"):
            collections.Sequence = collections.abc.Sequence      # removed from `collections` in Python 3.10
        if "easydict" not in sys.modules:
            _mod("easydict", EasyDict=AttrDict)
        for name in ("cv2", "matplotlib", "matplotlib.pyplot"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    _mod(name, use=lambda *a, **k: None)
        from tools.config.config import config
        from tools.dataset.cityscapes_vps import CityscapesVps
        config.dataset.num_seg_classes, config.dataset.num_classes = 19, 9
        _CACHE["tools_cls"] = CityscapesVps
    return _CACHE["tools_cls"]


def load_config(**overrides):
    cfg = runpy.run_path(os.path.join(REFERENCE_ROOT, CONFIG))
    cfg = to_attr({k: v for k, v in cfg.items() if not k.startswith("__")})
    cfg.model.pretrained = None
    for k, v in overrides.items():
        node = cfg.model
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg


def build_model(seed=0, **overrides):
    """build_detector(cfg.model) from the unchanged config, random init with ``seed``."""
    models = import_reference()
    cfg = load_config(**overrides)
    torch.manual_seed(seed)
    model = models.build_detector(cfg.model, train_cfg=None, test_cfg=cfg.test_cfg)
    return model.eval(), cfg
