"""TEST INFRASTRUCTURE — BASELINE.json configs[0] through the UNMODIFIED reference on CPU (this container only):
models/mm_late.MM_Model with real random-init BERT-base + ViT-B/16 towers (shims of oracle/ref_shims.py), trained by the
reference's own MMLate_Model.train loop (mm_late.py:416-532: AdamW over utils.get_optimizer_params, prepare_itm_inputs on the
numpy stream, the loss mix of :473-487) for three steps on the synthetic batches of oracle/config1_common.py, then its own eval.
Writes tests/golden/config1_losses.json (per-step training losses, the validation loss, trained logit_scale).

    python oracle/make_config1_golden.py
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from oracle import config1_common as K  # noqa: E402
from oracle.ref_shims import load_reference  # noqa: E402


def main():
    ns = load_reference(small_encoders=False)
    K.seed_all()
    model = ns.mm_late.MM_Model(K.C, "bernice", "vit", 0.0, fusion_name=K.FUSION)
    K.reinit_(model)
    train, val = K.synthetic_batches()
    M = ns.mm_late.MMLate_Model
    tr = types.SimpleNamespace(model=model, cnn=False, use_tim_loss=True, use_clip_loss=True, use_iadds_loss=False,
                               use_loss_correction=False, multilabel=False, softmax=nn.Softmax(dim=1), sigmoid=nn.Sigmoid(),
                               beta_itc=K.BETA_ITC, beta_itm=K.BETA_ITM, beta_iadds=0.1, num_labels=K.C)
    tr.prepare_itm_inputs = lambda ids, mask: M.prepare_itm_inputs(tr, ids, mask)
    tr.eval = lambda dl, loss_fn, tim_loss_fn=None, iadds_loss_fn=None: M.eval(tr, dl, loss_fn, tim_loss_fn=tim_loss_fn,
                                                                                iadds_loss_fn=iadds_loss_fn)
    K.seed_all()      # the ITM stream starts here on both sides
    with K.LossRecorder() as rec:
        M.train(tr, train, val, 1, nn.CrossEntropyLoss(), K.LR, K.WD, tim_loss_fn=nn.CrossEntropyLoss())
    model.eval()
    K.seed_all(123)
    res = M.eval(tr, val, nn.CrossEntropyLoss(), tim_loss_fn=nn.CrossEntropyLoss())
    out = {"train_losses": rec.losses, "val_loss": float(res["loss"]),
           "val_predictions": [int(x) for x in torch.as_tensor(res["predictions"]).reshape(-1)],
           "logit_scale": float(model.dual_encoder.logit_scale),
           "config": {"seed": K.SEED, "B": K.B, "L": K.L, "C": K.C, "fusion": K.FUSION, "lr": K.LR, "weight_decay": K.WD,
                      "towers": "BertModel(BertConfig()) + ViTModel(ViTConfig()), random init, dropout 0"}}
    path = os.path.join(ROOT, "tests", "golden", "config1_losses.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
