"""CLI mirror of the reference's models/run_mm_late.py: identical flag set (run_mm_late.py:20-43), identical seeding
(:48-49 — torch.manual_seed + np.random.seed, which drives the ITM negatives), identical output file naming (:91-96),
with the model replaced by tic_b200.mm_late.MMLate_Model (B200 kernels behind the same API).

    python -m tic_b200.run_mm_late --txt_model_name bernice --img_model_name vit --fusion_name attention \\
        --task 2 --epochs 7 --seed 40 --use_clip_loss --use_tim_loss

Data loading is not part of this path: `load_data` delegates to the reference's own `datasets.MM_Dataset` /
`utils.prepare_data`, so the reference's `models/` directory must be importable (PYTHONPATH) for a real run.
"""
import argparse
import logging

logging.basicConfig(format="%(asctime)s - %(message)s", datefmt="%Y-%m-%d %H:%M:%S", level=logging.INFO)
logger = logging.getLogger(__name__)


def build_parser():
    parser = argparse.ArgumentParser(description="run late fusion models")
    parser.add_argument("--txt_model_name", type=str, choices=["bert", "bernice", "bertweet", "roberta"], help="model name")
    parser.add_argument("--img_model_name", type=str, choices=["vit", "beit", "deit", "resnet50", "resnet152"], help="model name")
    parser.add_argument("--fusion_name", type=str, choices=["xatt", "concat", "attention", "concat_cnn", "aspect-att", "gmu"],
                        help="fusion method")
    parser.add_argument("--use_clip_loss", action="store_true", help="use contrastive Loss")
    parser.add_argument("--use_tim_loss", action="store_true", help="use TIM Loss")
    parser.add_argument("--use_iadds_loss", action="store_true", help="use image-adds loss")
    parser.add_argument("--beta_iadds", type=float, default=0.1, help="hyperparameter for iadds loss")
    parser.add_argument("--beta_itc", type=float, default=0.1, help="hyperparameter for itc loss")
    parser.add_argument("--beta_itm", type=float, default=0.1, help="hyperparameter for itm loss")
    parser.add_argument("--use_loss_correction", action="store_true", help="use Loss correction (only for binary cases)")
    parser.add_argument("--task", type=int, choices=[0, 1, 2, 3, 4, 5, 6], help="task to run")
    parser.add_argument("--epochs", type=int, default=2, help="number of epochs")
    parser.add_argument("--weight_decay", type=float, default=0.00025, help="weight decay param")
    parser.add_argument("--lr", type=float, default=1e-5, help="learning rate param")
    parser.add_argument("--dropout", type=float, default=0.05, help="dropout param")
    parser.add_argument("--seed", type=int, default=30, help="manual seed")
    parser.add_argument("--nsamples", type=int, default=-1, help="number of training samples")
    parser.add_argument("--testing", action="store_true", help="testing sample")
    parser.add_argument("--eval_txt_test", action="store_true", help="eval txt test")
    parser.add_argument("--save_model", action="store_true", help="save model")
    parser.add_argument("--load_saved_model", action="store_true", help="load saved model")
    parser.add_argument("--save_preds", action="store_true", help="eval test")
    parser.add_argument("--use_saved_features", action="store_true", help="use preprocessed features")
    # not in the reference: where the ITM coin flips / picks are drawn.  "numpy" replays the reference's global numpy stream
    # (seed-exact with the reference CLI); "device" draws and applies them on the GPU in one kernel
    parser.add_argument("--itm_rng", type=str, choices=["numpy", "device"], default="numpy", help="ITM sampling stream")
    return parser


def output_names(args, cfg, results_dir):
    """run_mm_late.py:87-96."""
    nsamples_str = "" if args.nsamples == -1 else "N" + str(args.nsamples) + "_"
    stem = "{}-{}-{}_task{}_seed{}_{}_{}".format(args.txt_model_name, args.img_model_name, args.fusion_name, args.task,
                                                 args.seed, cfg.loss_str, nsamples_str)
    model_path = results_dir + stem + "net.pth" if (args.save_model or args.load_saved_model) else None
    return model_path, results_dir + stem + "metrics_val.csv", results_dir + stem + "metrics_test.csv"


def aux_names(args, cfg, results_dir):
    """run_mm_late.py:124-126,148-162,178-187: prediction / metric files of --save_preds, --eval_txt_test, --load_saved_model."""
    nsamples_str = "" if args.nsamples == -1 else "N" + str(args.nsamples) + "_"
    stem = results_dir + "{}-{}-{}_task{}_seed{}_{}_{}".format(args.txt_model_name, args.img_model_name, args.fusion_name,
                                                                  args.task, args.seed, cfg.loss_str, nsamples_str)
    return {"preds": stem + "preds.csv", "preds_txt": stem + "preds_txt.csv", "metrics_txt": stem + "metrics_txt.csv",
            "preds_lm": stem + "preds_lm.csv", "metrics_lm": stem + "metrics_lm.csv"}


def save_predictions(predictions, filename):
    """run_mm_late.py:119-123: data_id / label / prediction columns."""
    import pandas as pd
    pd.DataFrame(data={"data_id": predictions["data_id"].tolist(), "label": predictions["labels"].tolist(),
                       "prediction": predictions["predictions"].tolist()}).to_csv(filename, index=False)
    logger.info("{} saved".format(filename))


def save_metrics(predictions, num_labels, filename):
    """run_mm_late.py:141-152,179-187: utils.compute_metrics -> CSV (the confusion matrix and the six scores are computed on the
    device).  The reference passes `multilabel=` to a function whose keyword is `multi_label` and raises on the
    --load_saved_model path; the call here is the intended one."""
    import pandas as pd
    from .utils import compute_metrics
    pd.DataFrame(compute_metrics(predictions, num_labels)).to_csv(filename, index=False)
    logger.info("{} saved".format(filename))


def main(argv=None):
    import numpy as np
    import torch
    import torch.nn as nn

    from .config import Config, results_dir_mm_late
    from .mm_late import MMLate_Model

    args = build_parser().parse_args(argv)
    torch.manual_seed(args.seed)     # run_mm_late.py:48
    np.random.seed(args.seed)        # run_mm_late.py:49 (drives prepare_itm_inputs)
    results_dir = results_dir_mm_late + ("testing/" if args.testing else "")
    logger.info("Model: {}-{}, Task: {}, Fusion: {}, Testing: {}, ITC Loss: {}, TIM Loss: {}, beta_itc: {}, beta_itm: {}, "
                "NSamples: {}, seed: {}".format(args.txt_model_name, args.img_model_name, args.task, args.fusion_name,
                                               args.testing, args.use_clip_loss, args.use_tim_loss, args.beta_itc,
                                               args.beta_itm, args.nsamples, args.seed))
    cfg = Config(args)
    mm_model = MMLate_Model(cfg, args.txt_model_name, args.img_model_name, args.fusion_name, multilabel=cfg.multilabel,
                            itm_rng=args.itm_rng)
    train_loader, val_loader, test_loader, weight, txt_te_loader = mm_model.load_data(
        cfg.data, cfg.img_fmt, testing=args.testing, nsamples=args.nsamples, saved_features=args.use_saved_features,
        task_name=cfg.task_name, eval_txt_test=args.eval_txt_test)
    weight = weight.to(mm_model.device) if weight is not None else None
    loss_fn = nn.CrossEntropyLoss(weight=weight) if not cfg.multilabel else nn.BCEWithLogitsLoss(pos_weight=weight)
    tim_loss_fn = nn.CrossEntropyLoss() if cfg.use_tim_loss else None
    model_path, val_filename, te_filename = output_names(args, cfg, results_dir)
    if not args.load_saved_model:
        logger.info("Training")
        mm_model.train(train_loader, val_loader, args.epochs, loss_fn, cfg.lr, cfg.weight_decay, tim_loss_fn=tim_loss_fn,
                       te_dataloader=test_loader, model_path=model_path, val_filename=val_filename, te_filename=te_filename)
        names = aux_names(args, cfg, results_dir)
        if args.save_preds:                                       # run_mm_late.py:117-127
            save_predictions(mm_model.eval(test_loader, loss_fn, tim_loss_fn=tim_loss_fn), names["preds"])
        if args.eval_txt_test and txt_te_loader is not None:      # run_mm_late.py:128-153
            logger.info("Evaluate and compute metrics (txt test)")
            predictions = mm_model.eval(txt_te_loader, loss_fn, tim_loss_fn=tim_loss_fn)
            save_predictions(predictions, names["preds_txt"])
            save_metrics(predictions, cfg.num_labels, names["metrics_txt"])
    else:                                                         # run_mm_late.py:156-187
        mm_model.load_saved_model(model_path)
        logger.info("Evaluate and compute metrics (test)")
        names = aux_names(args, cfg, results_dir)
        predictions = mm_model.eval(test_loader, loss_fn, tim_loss_fn=tim_loss_fn)
        save_predictions(predictions, names["preds_lm"])
        save_metrics(predictions, cfg.num_labels, names["metrics_lm"])
    logger.info("Done!")


if __name__ == "__main__":
    main()
