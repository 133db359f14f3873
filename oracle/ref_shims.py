"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (/root/reference/models) under the four shims of
SURVEY.md §8(c) so its own functions can pin the restatement in oracle/restatement.py and generate the
golden vectors under tests/golden/.  Nothing here is shipped or imported by the product package; the
reference tree does not exist on the GPU box, so only `oracle/make_golden.py` and the `reference`-marked
CPU tests (skipped when the tree is absent) call this.

Shims (none touches /root/reference):
  (i)   stub `torchmetrics` (+ `.classification`), imported by models/utils.py:10-14 but not installed;
  (ii)  `config.T`, imported by models/utils.py:16 yet never defined by models/config.py (reference bug);
  (iii) `transformers.ViTFeatureExtractor` (models/mm_late.py:10) was removed in transformers 5.x;
  (iv)  `VisionTextDualEncoderModel.from_vision_text_pretrained` (models/mm_late.py:59-61) needs local weight
        directories that do not exist: build random-init ViT-B/16 + BERT-base configs instead.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("TIC_REFERENCE_ROOT", "/root/reference")
REF_MODELS = os.path.join(REF_ROOT, "models")
REF_PREPROC = os.path.join(REF_ROOT, "preprocessing")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_MODELS, "mm_late.py"))


_loaded = {}


def load_reference(small_encoders: bool = True):
    """Returns a namespace with the reference modules `utils`, `mm_late`, `config`.

    small_encoders=True builds 2-layer encoders (hidden 768 is kept: the head hard-codes it,
    models/config.py:82-84) so a forward/backward takes milliseconds on CPU.
    """
    key = bool(small_encoders)
    if key in _loaded:
        return _loaded[key]
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    import torch  # noqa: F401
    import transformers

    # (i) torchmetrics stub
    if "torchmetrics" not in sys.modules:
        tm = types.ModuleType("torchmetrics")
        tmc = types.ModuleType("torchmetrics.classification")
        for name in ("F1Score", "Precision", "Recall", "Accuracy", "MultilabelF1Score", "MultilabelPrecision",
                     "MultilabelRecall", "MultilabelAccuracy"):
            cls = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: 0.0})
            setattr(tm, name, cls)
            setattr(tmc, name, cls)
        tm.classification = tmc
        sys.modules["torchmetrics"] = tm
        sys.modules["torchmetrics.classification"] = tmc

    # (iv) random-init dual encoder
    from transformers import BertConfig, BertModel, ViTConfig, ViTModel, VisionTextDualEncoderConfig
    from transformers import VisionTextDualEncoderModel

    def _from_vision_text_pretrained(cls_or_self=None, *args, **kwargs):
        nl = 2 if small_encoders else 12
        vcfg = ViTConfig(num_hidden_layers=nl, image_size=32 if small_encoders else 224)
        tcfg = BertConfig(num_hidden_layers=nl)
        cfg = VisionTextDualEncoderConfig.from_vision_text_configs(vcfg, tcfg)
        return VisionTextDualEncoderModel(config=cfg, vision_model=ViTModel(vcfg), text_model=BertModel(tcfg))

    VisionTextDualEncoderModel.from_vision_text_pretrained = classmethod(
        lambda cls, *a, **k: _from_vision_text_pretrained(cls, *a, **k))

    saved_path = list(sys.path)
    saved_cwd = os.getcwd()
    saved_mods = {k: sys.modules.get(k) for k in ("config", "utils", "datasets", "mm_late", "text_processing")}
    try:
        os.chdir(REF_MODELS)
        sys.path.insert(0, REF_PREPROC)
        sys.path.insert(0, REF_MODELS)
        for k in saved_mods:
            sys.modules.pop(k, None)
        import config as ref_config  # noqa: E402

        ref_config.T = [[0.9, 0.1], [0.1, 0.9]]  # (ii)
        import utils as ref_utils  # noqa: E402
        # (iii) ViTFeatureExtractor: set on the module object `from transformers import ...` resolves against,
        # after every other transformers import (the lazy module re-registers itself).
        from transformers import ViTImageProcessor
        for mod in {id(m): m for m in (sys.modules["transformers"], __import__("transformers", fromlist=["x"]))}.values():
            if "ViTFeatureExtractor" not in mod.__dict__:
                mod.__dict__["ViTFeatureExtractor"] = ViTImageProcessor
        import mm_late as ref_mm_late  # noqa: E402
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
    ns = types.SimpleNamespace(config=ref_config, utils=ref_utils, mm_late=ref_mm_late)
    # leave the reference modules importable only through `ns`
    for k, v in saved_mods.items():
        if v is not None:
            sys.modules[k] = v
        else:
            sys.modules.pop(k, None)
    _loaded[key] = ns
    return ns


def load_reference_early():
    """Adds `ns.mm_early` (models/mm_early.py) to the namespace of load_reference().  Extra shim (v): the module imports
    `lxmert_scripts.{modeling_frcnn,utils,processing_image}` (mm_early.py:10-12), a package that is NOT in the reference
    tree: stub modules with placeholder classes — only the encoder-free tail is ever called here
    (ViLT.get_logits_per_text :96-103, MMEarly_Model.prepare_itm_inputs :262-293)."""
    ns = load_reference()
    if getattr(ns, "mm_early", None) is not None:
        return ns
    stubs = {"lxmert_scripts": {}, "lxmert_scripts.modeling_frcnn": {"GeneralizedRCNN": type("GeneralizedRCNN", (), {})},
             "lxmert_scripts.utils": {"Config": type("Config", (), {})},
             "lxmert_scripts.processing_image": {"Preprocess": type("Preprocess", (), {})}}
    added = []
    for name, attrs in stubs.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            added.append(name)
    saved_path, saved_cwd = list(sys.path), os.getcwd()
    names = ("config", "utils", "datasets", "mm_early", "text_processing")
    saved_mods = {k: sys.modules.get(k) for k in names}
    try:
        os.chdir(REF_MODELS)
        sys.path.insert(0, REF_PREPROC)
        sys.path.insert(0, REF_MODELS)
        for k in names:
            sys.modules.pop(k, None)
        sys.modules["config"], sys.modules["utils"] = ns.config, ns.utils     # the already-shimmed modules (config.T, torchmetrics)
        import mm_early as ref_mm_early  # noqa: E402
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
        for name in added:
            sys.modules.pop(name, None)
    ns.mm_early = ref_mm_early
    return ns


def ref_early_logits(ns, text_embeds, image_embeds, logit_scale):
    """models/mm_early.py:96-103 (ViLT.get_logits_per_text), called unbound on a dummy self that owns `logit_scale`."""
    return ns.mm_early.ViLT.get_logits_per_text(types.SimpleNamespace(logit_scale=logit_scale), text_embeds, image_embeds)


def ref_early_prepare_itm_inputs(ns, ids, mask, token_type_ids):
    """models/mm_early.py:262-293, called unbound on a dummy self. Consumes np.random global state."""
    return ns.mm_early.MMEarly_Model.prepare_itm_inputs(types.SimpleNamespace(), ids, mask, token_type_ids)


def ref_eval(ns, outputs, labels, loss_fn):
    """models/mm_late.py:534-638 (MMLate_Model.eval) run UNBOUND on a dummy trainer whose model replays the given per-batch
    logits (no encoders, no aux losses): returns the reference's {data_id, loss, predictions, labels} for the eval bookkeeping."""
    import torch
    import torch.nn as nn
    it = iter(outputs)

    class _Model:
        def eval(self):
            return self

        def __call__(self, ids, mask, pixel_values, tim_inputs=None, iadds_task=False):
            return next(it), None, None, None, None

    dummy = types.SimpleNamespace(model=_Model(), cnn=False, use_tim_loss=False, use_clip_loss=False, use_iadds_loss=False,
                                  use_loss_correction=False, multilabel=False, softmax=nn.Softmax(dim=1), sigmoid=nn.Sigmoid(),
                                  beta_itc=0.0, beta_itm=0.0, beta_iadds=0.0)
    batches, n0 = [], 0
    for out, lab in zip(outputs, labels):
        B = out.shape[0]
        batches.append({"input_ids": torch.zeros(B, 1, 4, dtype=torch.long), "attention_mask": torch.ones(B, 1, 4, dtype=torch.long),
                        "pixel_values": torch.zeros(B, 1, 3, 2, 2), "labels": lab, "data_id": torch.arange(n0, n0 + B)})
        n0 += B
    return ns.mm_late.MMLate_Model.eval(dummy, batches, loss_fn)


def ref_prepare_itm_inputs(ns, ids, mask):
    """models/mm_late.py:389-414, called unbound on a dummy self (it uses no attributes). Consumes np.random global state."""
    dummy = types.SimpleNamespace()
    return ns.mm_late.MMLate_Model.prepare_itm_inputs(dummy, ids, mask)
