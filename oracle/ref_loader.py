"""TEST INFRASTRUCTURE ONLY — executes the reference's OWN Python source for the hot path, in place, from
/root/reference (never copied), under oracle/mmcv_stub.py. Used in THIS container to
  (1) generate the golden vectors committed under tests/golden/ (oracle/gen_golden.py), and
  (2) validate the travelling restatement oracle/fcos_oracle.py.
/root/reference does not exist on the GPU box: nothing that runs there may import this module.

`import mmdet` itself is impossible (mmcv, pycocotools, terminaltables missing — SURVEY.md §8c), so the mmdet
package tree is materialised as empty package shells whose __path__ points at the real directories; the hot
path's files are then imported normally (relative imports resolve against the real tree) while the heavy
package __init__ files are never run.
"""
import importlib
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("DSLB_REFERENCE_ROOT", "/root/reference")

_SHELLS = [
    "mmdet", "mmdet.core", "mmdet.core.bbox", "mmdet.core.bbox.iou_calculators", "mmdet.core.utils",
    "mmdet.core.mask", "mmdet.core.post_processing", "mmdet.core.export", "mmdet.core.visualization",
    "mmdet.models", "mmdet.models.losses", "mmdet.models.dense_heads", "mmdet.models.utils",
    "mmdet.models.backbones", "mmdet.models.necks", "mmdet.models.detectors", "mmdet.utils",
]


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mmdet"))


def _shell(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF_ROOT, *name.split("."))]
    m.__package__ = name
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


_loaded = None


def load():
    """Returns a namespace with the reference classes/functions of the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    from oracle import mmcv_stub
    mmcv_stub.install()
    warnings.filterwarnings("ignore", message=".*torch.meshgrid.*")
    for s in _SHELLS:
        _shell(s)
    imp = importlib.import_module

    # leaf utilities first, then publish the names `from mmdet.core import ...` expects
    core = sys.modules["mmdet.core"]
    masks = types.ModuleType("mmdet.core.mask.structures")
    masks.BitmapMasks = type("BitmapMasks", (), {})
    masks.PolygonMasks = type("PolygonMasks", (), {})
    sys.modules["mmdet.core.mask.structures"] = masks
    sys.modules["mmdet.core.visualization"].imshow_det_bboxes = None
    sys.modules["mmdet.utils"].get_root_logger = lambda *a, **k: __import__("logging").getLogger("mmdet")

    iou = imp("mmdet.core.bbox.iou_calculators.iou2d_calculator")
    sys.modules["mmdet.core.bbox.iou_calculators"].bbox_overlaps = iou.bbox_overlaps
    tr = imp("mmdet.core.bbox.transforms")
    du = imp("mmdet.core.utils.dist_utils")
    misc = imp("mmdet.core.utils.misc")
    onnx_helper = imp("mmdet.core.export.onnx_helper")
    sys.modules["mmdet.core.export"].get_k_for_topk = onnx_helper.get_k_for_topk
    core.bbox_overlaps = iou.bbox_overlaps
    core.distance2bbox = tr.distance2bbox
    core.bbox2result = tr.bbox2result
    core.bbox_mapping_back = tr.bbox_mapping_back
    core.merge_aug_proposals = None
    core.reduce_mean = du.reduce_mean
    core.multi_apply = misc.multi_apply
    nms = imp("mmdet.core.post_processing.bbox_nms")
    core.multiclass_nms = nms.multiclass_nms

    builder = imp("mmdet.models.builder")
    focal = imp("mmdet.models.losses.focal_loss")
    ioul = imp("mmdet.models.losses.iou_loss")
    ce = imp("mmdet.models.losses.cross_entropy_loss")
    res_layer = imp("mmdet.models.utils.res_layer")
    sys.modules["mmdet.models.utils"].ResLayer = res_layer.ResLayer
    resnet = imp("mmdet.models.backbones.resnet")
    fpn = imp("mmdet.models.necks.fpn")
    fcos_head = imp("mmdet.models.dense_heads.fcos_head")
    det_base = imp("mmdet.models.detectors.base")
    single_stage = imp("mmdet.models.detectors.single_stage")
    fcos = imp("mmdet.models.detectors.fcos")

    ns = types.SimpleNamespace(
        builder=builder, bbox_overlaps=iou.bbox_overlaps, distance2bbox=tr.distance2bbox,
        bbox2result=tr.bbox2result, reduce_mean=du.reduce_mean, multi_apply=misc.multi_apply,
        multiclass_nms=nms.multiclass_nms, FocalLoss=focal.FocalLoss,
        py_sigmoid_focal_loss=focal.py_sigmoid_focal_loss, GIoULoss=ioul.GIoULoss,
        CrossEntropyLoss=ce.CrossEntropyLoss, ResNet=resnet.ResNet, FPN=fpn.FPN, FCOSHead=fcos_head.FCOSHead,
        FCOS=fcos.FCOS, BaseDetector=det_base.BaseDetector, SingleStageDetector=single_stage.SingleStageDetector,
        fcos_head_module=fcos_head)
    _loaded = ns
    return ns


def load_rla():
    """RLA_ResNet (mmdet/models/backbones/resnet_rla.py:140-400), the backbone of the shipped DSL config
    (configs/fcos_semi/RLA_*.py:3-13), imported from the reference tree after load()."""
    ns = load()
    name = "mmdet.models.backbones.resnet_rla"
    if name not in sys.modules:
        # the module registers itself without force=True: drop a plugin class that already answers to the key
        ns.builder.BACKBONES._module_dict.pop("RLA_ResNet", None)
    return importlib.import_module(name).RLA_ResNet


def load_hook_functions():
    """parse_det_results / adathres from mmdet/runner/hooks/unlabel_pred_hook.py, executed without importing the
    module's heavy dependencies: only the two pure-Python function bodies are compiled from the source file."""
    import ast
    path = os.path.join(REF_ROOT, "mmdet/runner/hooks/unlabel_pred_hook.py")
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("parse_det_results", "adathres")]
    mod = ast.Module(body=keep, type_ignores=[])
    glb = {"os": os, "json": __import__("json"), "math": __import__("math")}
    exec(compile(mod, path, "exec"), glb)
    return glb["parse_det_results"], glb["adathres"]


def load_hook_chain():
    """save_results2file (+ gen_save_json_dict / parse_det_results / create_dir) from
    mmdet/runner/hooks/unlabel_pred_hook.py:20-175 compiled from the reference source file, with mmcv.ops.nms answered
    by the stub's restatement (torchvision nms: same IoU > thr rule, offset 0). Used only to generate golden vectors."""
    import ast
    import numpy as np
    from oracle import mmcv_stub
    path = os.path.join(REF_ROOT, "mmdet/runner/hooks/unlabel_pred_hook.py")
    tree = ast.parse(open(path).read())
    names = ("parse_det_results", "gen_save_json_dict", "create_dir", "save_results2file")
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    glb = {"os": os, "json": __import__("json"), "np": np, "nms": mmcv_stub.nms}
    exec(compile(mod, path, "exec"), glb)
    return glb["save_results2file"]
