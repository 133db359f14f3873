"""Drop-in mirror of the reference's models/mm_late.py for ViT-family image encoders: same class names, constructor and
method signatures, state-dict keys and return values, with everything after the HuggingFace encoders executed by the
hand-written sm_100a kernels of libtic_b200.so (no torch fallback: a CUDA device and the built library are required).

    MM_Model(num_labels, txt_model_name, img_model_name, dropout, fusion_name='concat')        mm_late.py:50-193
        .forward(ids, mask, pixel_values, tim_inputs=None, iadds_task=False)
            -> (out_cls, logits_per_text, out_tim, out_iadds, mm_features)
        .mm_fusion(x_t, x_v, x_v_pool=None, x_t_pool=None)                                      mm_late.py:91-144
    Scaled_Dot_Product_Attention().forward(Q, K, V, scale=None) -> (context, raw_scores)        mm_late.py:195-210
    MMLate_Model(config, txt_model_name, img_model_name, fusion_name, multilabel=False)         mm_late.py:298-739
        .prepare_itm_inputs(ids, mask) / .train(...) / .eval(...) / .load_saved_model(...)

What differs from the reference, by design:
  * the HF dual encoder is asked for the two towers' outputs only; projection, L2-normalise, the similarity GEMM and the
    (discarded) internal clip_loss of VisionTextDualEncoderModel.forward (HF :261-278) are NOT run by HF — the head does
    them (SURVEY.md K4');
  * the ITM branch does not run a second encoder pass (mm_late.py:170-175): the encoders are per-sample, so
    text_model(ids[src]) == text_model(ids)[src]; `tim_inputs` may carry the source rows as a third element
    (tim_ids, tim_mask, src_idx) — prepare_itm_inputs returns them — otherwise they are recovered by matching rows;
  * `attention` fusion evaluates the exact CLS-row collapse (only ctx[:,0,:] reaches the output, mm_late.py:111).
"""
import logging
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import capi
from .config import MODEL_DIR_DICT, fixed_feat_size, img_feat_size, txt_feat_size
from .plan import HeadPlan
from .utils import clip_loss, get_optimizer_params

logger = logging.getLogger(__name__)

HEAD_PARAM_ORDER = ["dual_encoder.text_projection.weight", "dual_encoder.visual_projection.weight", "dual_encoder.logit_scale",
                    "fc_Q.weight", "fc_Q.bias", "fc_K.weight", "fc_K.bias", "fc_V.weight", "fc_V.bias",
                    "aspectattention.weight", "aspectattention.bias", "linear_fusion.weight", "linear_fusion.bias",
                    "linear_cls.weight", "linear_cls.bias", "linear_tim.weight", "linear_tim.bias",
                    "linear_gmu_t.weight", "linear_gmu_t.bias", "linear_gmu_v.weight", "linear_gmu_v.bias"]
_GRAD_KEYS = {"dual_encoder.text_projection.weight": "dW_t", "dual_encoder.visual_projection.weight": "dW_v",
              "dual_encoder.logit_scale": "d_logit_scale", "fc_Q.weight": "dW_Q", "fc_Q.bias": "db_Q", "fc_K.weight": "dW_K",
              "fc_K.bias": "db_K", "fc_V.weight": "dW_V", "fc_V.bias": "db_V", "aspectattention.weight": "dw_a",
              "aspectattention.bias": "db_a", "linear_fusion.weight": "dW_f", "linear_fusion.bias": "db_f",
              "linear_cls.weight": "dW_cls", "linear_cls.bias": "db_cls", "linear_tim.weight": "dW_tim",
              "linear_tim.bias": "db_tim", "linear_gmu_t.weight": "dW_gt", "linear_gmu_t.bias": "db_gt",
              "linear_gmu_v.weight": "dW_gv", "linear_gmu_v.bias": "db_gv"}


class Scaled_Dot_Product_Attention(nn.Module):
    """mm_late.py:195-210, kept for API parity (the fused attention fusion does not call it: see fusion.cu)."""

    def forward(self, Q, K, V, scale=None):
        attention = torch.matmul(Q, K.permute(0, 2, 1))
        attention_scores = attention
        if scale:
            attention = attention * scale
        attention = F.softmax(attention, dim=-1)
        return torch.matmul(attention, V), attention_scores


class _HeadFn(torch.autograd.Function):
    """(x_t, x_v, t_pool, v_pool, *head params) -> (out_cls, logits_per_text, out_tim, mm_features) on a HeadPlan."""

    @staticmethod
    def forward(ctx, plan, aux, x_t, x_v, t_pool, v_pool, *params):
        dev = plan.dev
        bf = lambda t: t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()  # noqa: E731
        inp = {"x_t": bf(x_t), "x_v": bf(x_v), "t_pool": bf(t_pool), "v_pool": bf(v_pool)}
        inp.update(aux)
        # The plan reads the parameters through pointers (HeadPlan.bind_params): bound once per parameter storage; every
        # forward then refreshes the bf16 working copies and exp(logit_scale) on the device — no host synchronisation.
        key = tuple(p.data_ptr() for p in params)
        if getattr(plan, "_bound_key", None) != key:
            plan.bind_params(dict(zip(HEAD_PARAM_ORDER, params)), live=True)
            plan._bound_key = key
        out = plan.forward(inp)
        ctx.plan, ctx.inp, ctx.generation = plan, inp, plan.generation
        ctx.shapes = (x_t.shape, x_v.shape, x_t.dtype, t_pool.dtype, [p.dtype for p in params])
        out_tim = out["out_tim"].clone() if plan.use_itm else torch.zeros(0, device=dev)
        res = (out["out_cls"].clone(), out["logits_per_text"].clone(), out_tim, out["mm_features"].clone())
        ctx.mark_non_differentiable(res[3])   # mm_features feeds extract_features only (under no_grad in the reference)
        return res

    @staticmethod
    def backward(ctx, d_cls, d_logits, d_tim, _d_mm):
        plan = ctx.plan
        xt_shape, xv_shape, xt_dtype, tp_dtype, pdtypes = ctx.shapes
        # the activations live in the plan's buffers: a second forward on the same plan before this backward is an error
        o = plan.backward(ctx.inp, d_cls, d_logits, d_tim if plan.use_itm else None, generation=ctx.generation)
        d_x_t = None
        if "d_xt_cls" in o:   # x_t arrives as its CLS row only ([B,1,E], sliced by MM_Model.head)
            d_x_t = o["d_xt_cls"].to(xt_dtype).reshape(xt_shape).clone()
        d_tp = o["d_t_pool"].clone()
        if "d_t_pool_fusion" in o:
            d_tp = d_tp + o["d_t_pool_fusion"]
        grads = []
        for name, dt in zip(HEAD_PARAM_ORDER, pdtypes):
            g = o.get(_GRAD_KEYS[name])
            if g is None or (name.startswith("linear_tim") and not plan.use_itm):
                grads.append(None)
            elif name == "dual_encoder.logit_scale":
                grads.append(g.reshape(()).to(dt).clone())
            elif name == "aspectattention.weight":
                grads.append(g.reshape(1, -1).to(dt).clone())
            else:
                grads.append(g.to(dt).clone())
        # vision tower frozen (mm_late.py:67-69): no gradient for x_v / v_pool
        return (None, None, d_x_t, None, d_tp.to(tp_dtype), None) + tuple(grads)


class MM_Model(nn.Module):
    """mm_late.py:50-193. State-dict keys are the reference's (dual_encoder.*, fc_Q|fc_K|fc_V.*, aspectattention.*,
    linear_fusion.*, linear_cls.*, linear_tim.*, linear_iadds.*, linear_gmu_t|linear_gmu_v.*)."""

    def __init__(self, num_labels, txt_model_name, img_model_name, dropout, fusion_name="concat", dual_encoder=None):
        super().__init__()
        self.num_labels, self.fusion_name = num_labels, fusion_name
        self.txt_model_name, self.img_model_name = txt_model_name, img_model_name
        if dual_encoder is None:
            from transformers import VisionTextDualEncoderModel
            dual_encoder = VisionTextDualEncoderModel.from_vision_text_pretrained(MODEL_DIR_DICT[img_model_name],
                                                                                  MODEL_DIR_DICT[txt_model_name])
        self.dual_encoder = dual_encoder
        for name, param in self.dual_encoder.named_parameters():   # freeze vision (mm_late.py:67-69)
            if "vision" in name:
                param.requires_grad = False
        self.dropout = nn.Dropout(dropout)
        self.fc_Q = nn.Linear(txt_feat_size, fixed_feat_size)
        self.fc_K = nn.Linear(img_feat_size, fixed_feat_size)
        self.fc_V = nn.Linear(img_feat_size, fixed_feat_size)
        self.attention = Scaled_Dot_Product_Attention()
        self.aspectattention = nn.Linear(fixed_feat_size, 1)
        self.m = nn.Softmax(dim=1)
        self.linear_fusion = nn.Linear(fixed_feat_size * 2, fixed_feat_size)
        self.relu = nn.ReLU()
        self.linear_cls = nn.Linear(fixed_feat_size, self.num_labels)
        self.linear_tim = nn.Linear(fixed_feat_size, 2)
        self.linear_iadds = nn.Linear(fixed_feat_size, 2)
        self.z = nn.Sigmoid()
        self.tanh = nn.Tanh()
        self.linear_gmu_t = nn.Linear(fixed_feat_size, 2 * fixed_feat_size)
        self.linear_gmu_v = nn.Linear(fixed_feat_size, 2 * fixed_feat_size)
        self._plans = {}

    # ---------------------------------------------------------------- plumbing
    def _head_params(self):
        named = dict(self.named_parameters())
        return [named[n] for n in HEAD_PARAM_ORDER]

    def _plan(self, B, use_itm, Lv, dev):
        key = (B, bool(use_itm), Lv, str(dev))
        if key not in self._plans:
            if self.fusion_name not in ("concat", "attention", "aspect-att", "gmu"):
                # mm_late.py: mm_fusion falls through and returns None -> TypeError in dropout; xatt / concat_cnn are
                # undefined names in the reference (mm_late.py:44-45)
                raise TypeError("fusion_name %r is not implemented by mm_fusion (mm_late.py:91-144)" % self.fusion_name)
            P = self.dual_encoder.text_projection.weight.shape[0]
            self._plans[key] = HeadPlan(B, E=fixed_feat_size, P=P, C=self.num_labels, fusion=self.fusion_name, use_itc=True,
                                        use_itm=use_itm, Lv=Lv, materialize_logits=True, device=dev)
        return self._plans[key]

    def _encode(self, ids, mask, pixel_values):
        """The two untouched HF towers (mm_late.py:149-158 minus the HF-side ITC tail)."""
        v = self.dual_encoder.vision_model(pixel_values=pixel_values)
        t = self.dual_encoder.text_model(input_ids=ids, attention_mask=mask)
        return t.last_hidden_state, t.pooler_output, v.last_hidden_state, v.pooler_output

    @staticmethod
    def _source_rows(ids, tim_ids):
        """src[i] = row of `ids` that sits at position i of `tim_ids` (first match; the reference copies whole rows)."""
        eq = (tim_ids[:, None, :] == ids[None, :, :]).all(dim=2)
        return eq.float().argmax(dim=1).to(torch.int32)

    def head(self, x_t, x_v, x_t_pool, x_v_pool, tim_src=None):
        """Everything after the encoders: returns (out_cls, logits_per_text, out_tim, mm_features)."""
        if not x_t.is_cuda:
            raise capi.TicError("tic_b200.MM_Model needs CUDA tensors: this package has no CPU path")
        B = x_t.shape[0]
        use_itm = tim_src is not None
        plan = self._plan(B, use_itm, x_v.shape[1], x_t.device)
        aux = {}
        if use_itm:
            aux["src_idx"] = tim_src.to(device=x_t.device, dtype=torch.int32)
            aux["lbl_tim"] = torch.zeros(B, dtype=torch.int64, device=x_t.device)  # labels live in the caller's loss
        p = self.dropout.p
        if self.training and p > 0:
            aux["keep"] = (torch.rand(B, fixed_feat_size, device=x_t.device) >= p).to(torch.uint8)
            aux["keep_scale"] = 1.0 / (1.0 - p)
        # every fusion variant reads only the CLS row of x_t (mm_late.py:94,111,141): slice here, so that the cast to bf16 and
        # the gradient handed back to autograd are [B,1,E] instead of [B,Lt,E]
        out_cls, logits, out_tim, mm = _HeadFn.apply(plan, aux, x_t[:, :1], x_v, x_t_pool, x_v_pool, *self._head_params())
        return out_cls, logits, (out_tim if use_itm else None), mm

    # ---------------------------------------------------------------- reference surface
    def mm_fusion(self, x_t, x_v, x_v_pool=None, x_t_pool=None):
        """mm_late.py:91-144 (forward only, no autograd): the fused vector [B, 768] of the configured fusion."""
        if self.fusion_name == "aspect-att" and (x_v_pool is None or x_t_pool is None):
            raise TypeError("expected Tensor as element 0 in argument 0, but got NoneType")  # torch.stack((None, None))
        B, dev = x_t.shape[0], x_t.device
        zeros = torch.zeros(B, fixed_feat_size, device=dev)
        with torch.no_grad():
            out = self.head(x_t, x_v, x_t_pool if x_t_pool is not None else zeros,
                            x_v_pool if x_v_pool is not None else zeros)
        return out[3]

    def forward(self, ids, mask, pixel_values, tim_inputs=None, iadds_task=False):
        x_t, x_t_pool, x_v, x_v_pool = self._encode(ids, mask, pixel_values)
        tim_src = None
        if tim_inputs is not None:
            if self.fusion_name == "aspect-att":   # mm_late.py:181: mm_fusion without pools -> torch.stack((None, None))
                raise TypeError("expected Tensor as element 0 in argument 0, but got NoneType")
            tim_src = tim_inputs[2] if len(tim_inputs) > 2 else self._source_rows(ids, tim_inputs[0])
        out_cls, logits_per_text, out_tim, mm_features = self.head(x_t, x_v, x_t_pool, x_v_pool, tim_src)
        out_iadds = None
        if iadds_task:   # dead branch in the reference (Config forces use_iadds_loss False, config.py:65); library call
            out_iadds = self.linear_iadds(self.dropout(mm_features))
        return out_cls, logits_per_text, out_tim, out_iadds, mm_features


class MMLate_Model(object):
    """mm_late.py:298-739: trainer wrapper.  Data loading is out of scope of this path: pass
    DataLoaders yielding the reference's batch dict (input_ids, attention_mask, pixel_values, labels, data_id)."""

    def __init__(self, config, txt_model_name, img_model_name, fusion_name, multilabel=False, model=None, device=None,
                 itm_rng="numpy"):
        self.batch_size, self.num_labels, self.multilabel = config.batch_size, config.num_labels, multilabel
        self.use_clip_loss, self.beta_itc = config.use_clip_loss, config.beta_itc
        self.use_tim_loss, self.beta_itm = config.use_tim_loss, config.beta_itm
        self.use_iadds_loss, self.beta_iadds = config.use_iadds_loss, config.beta_iadds
        self.use_loss_correction = config.use_loss_correction
        self.txt_model_name, self.img_model_name, self.max_length = txt_model_name, img_model_name, config.max_length
        self.cnn = img_model_name in {"resnet50", "resnet152"}
        if self.cnn:
            raise NameError("name 'XATT' is not defined")  # mm_late.py:42-47: the CNN fusions are undefined in the reference
        self.device = torch.device(device if device is not None else "cuda:0")
        self.model = model if model is not None else MM_Model(self.num_labels, txt_model_name, img_model_name,
                                                              config.dropout, fusion_name=fusion_name)
        self.model.to(self.device)
        self.softmax, self.sigmoid = nn.Softmax(dim=1), nn.Sigmoid()
        # "numpy": ITM decisions replay the reference's global numpy stream (seed-exact CLI parity); "device": drawn and applied
        # on the GPU by one kernel, no host loop in the training step
        self.itm_rng = itm_rng
        self.last_eval = None

    def load_saved_model(self, model_path):
        self.model.load_state_dict(torch.load(model_path))

    def load_data(self, data, img_file_fmt, testing=False, nsamples=-1, saved_features=False, task_name=None,
                  eval_txt_test=False, compute_class_weights=True, random_labels=False):
        """mm_late.py:346-387.  Dataset construction is outside this path: it is delegated to
        the reference's own `utils.prepare_data` and `datasets.MM_Dataset`, which must be importable."""
        if eval_txt_test:
            # reference mm_late.py:372-387 builds a text-only test loader from a second CSV; dataset plumbing is outside this
            # path, and returning None here would make --eval_txt_test silently skip the preds_txt / metrics_txt files
            raise NotImplementedError("eval_txt_test: the text-only test loader (reference mm_late.py:372-387) is not built by "
                                      "this package — construct it with the reference's datasets.MM_Dataset and call eval() on it")
        try:
            from datasets import MM_Dataset          # reference models/datasets.py
            from utils import prepare_data           # reference models/utils.py
        except Exception as e:  # pragma: no cover
            raise ImportError("load_data needs the reference's models/ directory on PYTHONPATH (datasets.MM_Dataset, "
                              "utils.prepare_data); data loading is not part of the accelerated path") from e
        from torch.utils.data import DataLoader
        from transformers import AutoTokenizer, ViTImageProcessor, VisionTextDualEncoderProcessor
        tok = AutoTokenizer.from_pretrained(MODEL_DIR_DICT[self.txt_model_name])
        proc = VisionTextDualEncoderProcessor(ViTImageProcessor.from_pretrained(MODEL_DIR_DICT[self.img_model_name]), tok)
        train, y_tr, val, y_val, test, y_te, class_weights, image_adds = prepare_data(
            data, self.num_labels, testing=testing, nsamples=nsamples, compute_class_weights=compute_class_weights,
            random_labels=random_labels, load_image_adds=self.use_iadds_loss, multilabel=self.multilabel)
        mk = lambda df, y, key: MM_Dataset(df.tweet_id.values, df.text.values, y, proc, self.max_length,  # noqa: E731
                                           img_file_fmt=img_file_fmt, saved_features=saved_features, task_name=task_name,
                                           image_adds=image_adds[key])
        return (DataLoader(mk(train, y_tr, "train"), batch_size=self.batch_size, shuffle=True),
                DataLoader(mk(val, y_val, "val"), batch_size=self.batch_size, shuffle=False),
                DataLoader(mk(test, y_te, "test"), batch_size=self.batch_size, shuffle=False), class_weights, None)

    def prepare_itm_inputs(self, ids, mask, return_src=False, rng="numpy", generator=None):
        """mm_late.py:389-414.  rng="numpy" (default, seed-exact with the reference CLI): consumes the GLOBAL numpy stream
        exactly like the reference (coin, then pick, per row), then performs all row copies with the device gather kernel.
        rng="device" (SURVEY §8 f-2): no host loop and no host->device copy at all — the two uniforms per row come from
        torch's CUDA generator (`generator`, default the global one) and ONE kernel applies the same rule (swap iff
        u_coin < 0.5; uniform pick over the other B-1 rows) and gathers ids and mask.  Same distribution, different stream.
        Returns fresh tensors (tim_ids, tim_mask, lbl_tim), plus the source rows when return_src=True (so forward() need not
        re-derive them)."""
        if not ids.is_cuda:
            raise capi.TicError("prepare_itm_inputs needs CUDA tensors: this package has no CPU path")
        B = ids.shape[0]
        if rng == "device":
            dev = ids.device
            ids_c, mask_c = ids.contiguous(), mask.contiguous()
            if ids_c.dtype != mask_c.dtype or ids_c.shape != mask_c.shape:
                raise ValueError("ids and mask must share dtype and shape")
            u = torch.rand(2, B, device=dev, generator=generator)
            tim_ids, tim_mask = torch.empty_like(ids_c), torch.empty_like(mask_c)   # never aliases its inputs (:391-392)
            lbl_tim = torch.empty(B, dtype=torch.int64, device=dev)
            src_d = torch.empty(B, dtype=torch.int32, device=dev)
            capi.call("tic_itm_sample_gather", u[0].data_ptr(), u[1].data_ptr(), B, 0, None, 0, 0.0, None, ids_c.data_ptr(), mask_c.data_ptr(),
                      ids_c.stride(0) * ids_c.element_size(), tim_ids.data_ptr(), tim_mask.data_ptr(), lbl_tim.data_ptr(),
                      src_d.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return (tim_ids, tim_mask, lbl_tim, src_d) if return_src else (tim_ids, tim_mask, lbl_tim)
        if rng != "numpy":
            raise ValueError("rng must be 'numpy' or 'device'")
        swap, src = _decisions_from_numpy_stream(B)
        dev = ids.device
        ids_c, mask_c = ids.contiguous(), mask.contiguous()
        src_d = torch.from_numpy(src.astype(np.int32)).to(dev)
        tim_ids, tim_mask = torch.empty_like(ids_c), torch.empty_like(mask_c)   # never aliases its inputs (:391-392)
        st = torch.cuda.current_stream().cuda_stream
        for s_, d_ in ((ids_c, tim_ids), (mask_c, tim_mask)):
            capi.call("tic_gather_rows", s_.data_ptr(), s_.stride(0) * s_.element_size(), d_.data_ptr(),
                      d_.stride(0) * d_.element_size(), s_.shape[1] * s_.element_size(), src_d.data_ptr(), B, st)
        lbl_tim = torch.from_numpy((~swap).astype(np.int64)).to(dev)
        if return_src:
            return tim_ids, tim_mask, lbl_tim, src_d
        return tim_ids, tim_mask, lbl_tim

    def _batch(self, batch):
        dev = self.device
        ids = torch.squeeze(batch["input_ids"])
        mask = torch.squeeze(batch["attention_mask"])
        pixel_values = torch.squeeze(batch["pixel_values"])
        if pixel_values.dim() < 4:
            pixel_values, ids, mask = pixel_values.unsqueeze(0), ids.unsqueeze(0), mask.unsqueeze(0)
        return ids.to(dev), mask.to(dev), pixel_values.to(dev)

    def _loss(self, loss_fn, tim_loss_fn, output, label, logits_per_text, output_tim, lbl_tim):
        """mm_late.py:473-487 / :581-593."""
        if self.use_clip_loss and self.use_tim_loss:
            return (1 - (self.beta_itc + self.beta_itm)) * loss_fn(output, label) + self.beta_itc * clip_loss(logits_per_text) \
                + self.beta_itm * tim_loss_fn(output_tim, lbl_tim)
        if self.use_clip_loss:
            return (1 - self.beta_itc) * loss_fn(output, label) + self.beta_itc * clip_loss(logits_per_text)
        if self.use_tim_loss:
            return (1 - self.beta_itm) * loss_fn(output, label) + self.beta_itm * tim_loss_fn(output_tim, lbl_tim)
        return loss_fn(output, label)

    def train(self, dataloader, val_dataloader, epochs, loss_fn, lr, weight_decay, tim_loss_fn=None, iadds_loss_fn=None,
              te_dataloader=None, model_path=None, val_filename=None, te_filename=None):
        """mm_late.py:416-532 (AdamW over requires_grad params; per-batch forward, loss mix, backward, step)."""
        optimizer = torch.optim.AdamW(get_optimizer_params(self.model.named_parameters(), weight_decay, lr), lr=lr)
        res_val, res_te = [], []
        for epoch in range(epochs):
            self.model.train()
            for batch in dataloader:
                ids, mask, pixel_values = self._batch(batch)
                label = batch["labels"].to(self.device)
                optimizer.zero_grad()
                tim_inputs, lbl_tim = self._itm_inputs(ids, mask)
                output, logits_per_text, output_tim, _, _ = self.model(ids, mask, pixel_values, tim_inputs=tim_inputs,
                                                                       iadds_task=self.use_iadds_loss)
                label = label.type_as(output)
                loss = self._loss(loss_fn, tim_loss_fn, output, label, logits_per_text, output_tim, lbl_tim)
                loss.backward()
                optimizer.step()
                optimizer.zero_grad()
            write = epoch % 2 == 0 or epoch == epochs - 1          # mm_late.py:511,523: every 2nd epoch and the last one
            if val_dataloader is not None:
                r = self.eval(val_dataloader, loss_fn, tim_loss_fn=tim_loss_fn, iadds_loss_fn=iadds_loss_fn)
                r["epoch"] = epoch
                res_val.append(r)
                if val_filename is not None and write:
                    self._write_metrics(res_val, val_filename)
            if te_dataloader is not None:
                r = self.eval(te_dataloader, loss_fn, tim_loss_fn=tim_loss_fn, iadds_loss_fn=iadds_loss_fn)
                r["epoch"] = epoch
                res_te.append(r)
                if te_filename is not None and write:
                    self._write_metrics(res_te, te_filename)
        if model_path is not None:
            torch.save(self.model.state_dict(), model_path)
            logger.info("{} saved".format(model_path))
        return res_val, res_te

    def _itm_inputs(self, ids, mask):
        """(tim_inputs, lbl_tim) for one batch, or (None, None); the reference samples ITM negatives in eval, prediction and
        feature extraction too (mm_late.py:567,669,719)."""
        if not self.use_tim_loss:
            return None, None
        tim_ids, tim_mask, lbl_tim, src = self.prepare_itm_inputs(ids, mask, return_src=True, rng=self.itm_rng)
        return (tim_ids, tim_mask, src), lbl_tim

    def _write_metrics(self, results, filename):
        """mm_late.py:511-527: one column per epoch (utils.agg_metrics_val over config.metric_names) -> CSV."""
        import pandas as pd
        from .config import metric_names
        from .eval import agg_metrics_val
        pd.DataFrame(agg_metrics_val(results, metric_names, self.num_labels)).to_csv(filename, index=False)
        logger.info("{} saved!".format(filename))

    def eval(self, dataloader, loss_fn, tim_loss_fn=None, iadds_loss_fn=None):
        """mm_late.py:534-638 — returns {data_id, loss, predictions, labels}.  Predictions, targets, the loss sum and the
        confusion matrix are accumulated on the device (eval.EvalAccumulator): one host synchronisation per
        epoch instead of two per batch; `self.last_eval` keeps the accumulator (confusion matrix, device-side metrics)."""
        from .eval import EvalAccumulator
        if self.multilabel:
            raise NotImplementedError("multilabel evaluation belongs to the unreachable task 10 of the reference (config.py:10)")
        self.model.eval()
        try:
            capacity = len(dataloader.dataset)
        except (TypeError, AttributeError):
            capacity = len(dataloader) * self.batch_size
        acc = EvalAccumulator(self.num_labels, capacity, device=self.device)
        for batch in dataloader:
            ids, mask, pixel_values = self._batch(batch)
            label = batch["labels"].to(self.device)
            data_id = batch["data_id"].to(self.device)
            with torch.no_grad():
                tim_inputs, lbl_tim = self._itm_inputs(ids, mask)
                output, logits_per_text, output_tim, _, _ = self.model(ids, mask, pixel_values, tim_inputs=tim_inputs,
                                                                       iadds_task=self.use_iadds_loss)
                label = label.type_as(output)
                loss = self._loss(loss_fn, tim_loss_fn, output, label, logits_per_text, output_tim, lbl_tim)
            acc.update(output, label, loss=loss, data_id=data_id)
        res = acc.result()
        self.last_eval = acc
        logger.info("loss: %.4f acc: %.4f", res["loss"], res["accuracy"])
        return {"data_id": res.get("data_id"), "loss": res["loss"], "predictions": res["predictions"], "labels": res["labels"]}

    def compute_predictions(self, dataloader):
        """mm_late.py:640-701 — {data_id, predictions}.  (The reference unpacks 4 of the model's 5 outputs at :674 and would
        raise; SURVEY §9: keep the method, fix the unpack.)"""
        self.model.eval()
        predictions, data_ids = [], []
        for batch in dataloader:
            ids, mask, pixel_values = self._batch(batch)
            data_id = batch["data_id"].to(self.device)
            with torch.no_grad():
                tim_inputs, _ = self._itm_inputs(ids, mask)
                output, _, _, _, _ = self.model(ids, mask, pixel_values, tim_inputs=tim_inputs, iadds_task=self.use_iadds_loss)
            if self.multilabel:
                pred = torch.round(self.sigmoid(output))
            else:
                pred = torch.argmax(self.softmax(output), dim=1)
            predictions.append(pred)
            data_ids.append(data_id)
        return {"data_id": torch.cat(data_ids), "predictions": torch.cat(predictions)}

    def extract_features(self, dataloader):
        """mm_late.py:703-739 — (mm_features [N, 768], argmax labels [N])."""
        self.model.eval()
        features, labels = [], []
        for batch in dataloader:
            ids, mask, pixel_values = self._batch(batch)
            label = batch["labels"].to(self.device)
            with torch.no_grad():
                tim_inputs, _ = self._itm_inputs(ids, mask)
                _, _, _, _, mm_feats = self.model(ids, mask, pixel_values, tim_inputs=tim_inputs, iadds_task=self.use_iadds_loss)
            features.append(mm_feats)
            labels.append(torch.argmax(label, dim=1))
        return torch.cat(features), torch.cat(labels)


def _decisions_from_numpy_stream(B):
    """The reference's draws (mm_late.py:396-401) on the GLOBAL numpy stream, vectorised per row:
    np.random.choice([True, False]) then np.random.choice(list(set(range(B)) - {idx}))."""
    swap = np.zeros(B, dtype=bool)
    src = np.arange(B, dtype=np.int64)
    if B > 1:
        for idx in range(B):
            if np.random.choice([True, False]):
                swap[idx] = True
                src[idx] = np.random.choice(list(set(range(B)) - {idx}))
    return swap, src
