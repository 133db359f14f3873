"""TEST INFRASTRUCTURE — shared by oracle/make_config1_golden.py (unmodified reference on CPU, this container) and
tests/test_gpu_config1.py (tic_b200 on the B200): BASELINE.json configs[0] — "Ber-ViT-Att late fusion, task 2, --testing,
random-init weights" (README.md:35-38, models/run_mm_late.py:65-191) on REAL HuggingFace BERT-base + ViT-B/16 towers.

There are no weights, tokenizer files or tweets on either machine: the run uses synthetic token ids / pixel
tensors of the reference's batch layout (datasets.MM_Dataset: input_ids [B,1,L], attention_mask [B,1,L], pixel_values
[B,1,3,224,224], float one-hot labels [B,C], data_id [B]) and a deterministic per-parameter re-initialisation, so that both
sides hold bit-identical fp32 weights without depending on the order modules consume the torch RNG.  Every nn.Dropout is set
to p = 0 (towers and head): the reference's second encoder pass for the ITM pairs (mm_late.py:170-175) then equals the
row gather x_t[src] this implementation performs, and CPU / CUDA dropout streams cannot make the runs diverge."""
import hashlib

import numpy as np
import torch

SEED, B, L, C, N_TRAIN, LR, WD = 40, 8, 128, 4, 3, 1e-5, 0.00025
FUSION, BETA_ITC, BETA_ITM = "attention", 0.1, 0.1


def synthetic_batches(n_batches=N_TRAIN + 1, seed=SEED):
    g = torch.Generator().manual_seed(seed)
    out = []
    for k in range(n_batches):
        ids = torch.randint(1000, 28000, (B, 1, L), generator=g)
        ids[:, :, 0] = 101
        lens = torch.randint(8, L, (B,), generator=g)
        mask = (torch.arange(L)[None, None, :] < lens[:, None, None]).long()
        ids = ids * mask
        y = torch.randint(0, C, (B,), generator=g)
        out.append({"input_ids": ids, "attention_mask": mask, "pixel_values": torch.randn(B, 1, 3, 224, 224, generator=g),
                    "labels": torch.eye(C)[y], "data_id": torch.arange(k * B, (k + 1) * B)})
    return out[:-1], out[-1:]      # train batches, one validation batch


def reinit_(model, seed=SEED):
    """Deterministic weights keyed by PARAMETER NAME (the state-dict keys are the reference's on both sides)."""
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            h = int.from_bytes(hashlib.sha256(("%d:%s" % (seed, name)).encode()).digest()[:8], "little") % (2 ** 63)
            g = torch.Generator().manual_seed(h)
            v = torch.randn(p.shape, generator=g) * 0.02
            if name.endswith("logit_scale"):
                v = torch.full(p.shape, 2.6592)
            elif "LayerNorm.weight" in name or "layernorm" in name.lower() and name.endswith("weight"):
                v = 1.0 + v
            p.copy_(v.to(p.dtype))
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    return model


class LossRecorder:
    """records float(loss) at every loss.backward() of the training loop (the reference prints accuracy only)"""

    def __init__(self):
        self.losses, self._orig = [], None

    def __enter__(self):
        self._orig = torch.Tensor.backward
        rec = self

        def backward(t, *a, **k):
            if t.dim() == 0:
                rec.losses.append(float(t.detach()))
            return rec._orig(t, *a, **k)
        torch.Tensor.backward = backward
        return self

    def __exit__(self, *a):
        torch.Tensor.backward = self._orig


def seed_all(seed=SEED):
    torch.manual_seed(seed)      # run_mm_late.py:48
    np.random.seed(seed)         # run_mm_late.py:49 (drives prepare_itm_inputs)
