"""Mirror of the reference's models/config.py for the mm_late path: same `Config(args, ...)` attributes and the same
module-level constants (sizes, task tables, paths, model directories), so `run_mm_late.py` behaves identically.
Reference: models/config.py:1-77 (Config), :82-152 (constants)."""
import os

txt_feat_size = 768      # config.py:82
fixed_feat_size = 768    # config.py:83
img_feat_size = 768      # config.py:84
img_feat_size_cnn = 2048

TASKS = {0: "text_is_represented", 1: "image_adds", 2: "tir", 3: "mvsa", 4: "mhp", 5: "mic", 6: "msd"}
DATA_PATH = os.environ.get("TIC_DATA_PATH", "../data/")
PATH = {0: DATA_PATH + "data_key_imgtxt_random.csv", 1: DATA_PATH + "data_key_imgtxt_random.csv",
        2: DATA_PATH + "data_key_imgtxt_random.csv", 3: DATA_PATH + "data_key_mvsa.csv", 4: DATA_PATH + "data_key_mhp.csv",
        5: DATA_PATH + "data_key_mic.csv", 6: DATA_PATH + "data_key_msd.csv"}
IMG_FMT = {0: DATA_PATH + "text-image/T{}.jpg", 1: DATA_PATH + "text-image/T{}.jpg", 2: DATA_PATH + "text-image/T{}.jpg",
           3: DATA_PATH + "MVSA-Single/data/{}.jpg", 4: DATA_PATH + "MHP/Data/Images/{}.jpg",
           5: DATA_PATH + "MIC/spc_imgs_twitter/{}_1.jpg", 6: DATA_PATH + "MSD/dataset_image/{}.jpg"}
CLASSES = {2: ["image adds and text is represented", "image adds and text is not represented",
               "image does not add and text is represented", "image does not adds and text is not represented"],
           3: ["neutral", "positive", "negative"], 6: ["not sarcastic", "sarcastic"]}
EMPTY_IMG = DATA_PATH + "MIC/empty_image.png"
metric_names = ["f1_weighted", "f1_macro", "precision_weighted", "precision_macro", "recall_weighted", "recall_macro", "loss"]
RES_PATH = os.environ.get("TIC_RESULTS_PATH", "../results/")
results_dir_mm_late = RES_PATH + "mm_late/"
IMAGE_ADDS = results_dir_mm_late + "bernice-vit-attention_task{}_seed30_preds_lm.csv"
MODEL_DIR_DICT = {"bert": "../../../BERT-base/", "bertweet": "../../../BERTWEET-base/", "roberta": "../../../RoBERTa-base/",
                  "bernice": "../../../BERNICE/", "vit": "../../../ViT/", "beit": "../../../BEiT/", "deit": "../../../DEiT/",
                  "resnet50": "../../../ConvModels/resnet50-0676ba61.pth", "resnet152": "../../../ConvModels/resnet152-394f9c45.pth"}
# the reference imports T from config (models/utils.py:16) but never defines it: loss correction is deprecated (config.py:77)
T = [[0.9, 0.1], [0.1, 0.9]]


class Config(object):
    """models/config.py:1-77. `args` is the argparse namespace of run_mm_late.py; reads PATH[task] with pandas."""

    def __init__(self, args, model_name=None, multimodal=True, txt=False, data_key=None):
        import numpy as np
        self.multilabel = True if args.task in {10} else False
        self.column_names = ["tweet_id", "text", "label", "split"]
        if data_key is None:
            import pandas as pd
            data_key = pd.read_csv(PATH[args.task])
        if args.task < 2:
            self.data = data_key[["tweet_id", "text", TASKS[args.task], "split"]].rename(columns={TASKS[args.task]: "label"})
            self.num_labels, self.batch_size = 2, 8
        elif args.task == 2:
            data = data_key[["tweet_id", "text", "split"]].copy()
            df_labels = data_key[["image_adds_text_repr", "image_adds_text_notrepr", "image_notadds_text_repr",
                                  "image_notadds_text_notrepr"]].to_numpy()
            data["label"] = np.argmax(df_labels, axis=1)
            self.data = data[["tweet_id", "text", "label", "split"]]
            self.num_labels, self.batch_size = 4, 8
        elif args.task == 3:
            self.data, self.num_labels, self.batch_size = data_key[self.column_names], 3, 16
        elif args.task == 4:
            self.data, self.num_labels, self.batch_size = data_key[self.column_names], 4, 8
        elif args.task == 5:
            self.data = data_key[["id", "text", "label", "split"]].rename(columns={"id": "tweet_id"})
            self.num_labels, self.batch_size = 2, 16
        elif args.task == 6:
            self.data, self.num_labels, self.batch_size = data_key[self.column_names], 2, 16
        self.img_fmt = IMG_FMT[args.task]
        self.task_name = TASKS[args.task]
        self.classes = CLASSES[args.task] if args.task in CLASSES else None
        self.dropout, self.weight_decay, self.lr = args.dropout, args.weight_decay, args.lr
        self.max_length = 40 if (model_name is not None and model_name == "vilt") else 128
        if multimodal:
            self.use_clip_loss = args.use_clip_loss
            self.use_tim_loss = args.use_tim_loss
            self.use_iadds_loss = False  # deprecated (config.py:65): the flag is parsed and ignored
            self.beta_itc = args.beta_itc if self.use_clip_loss else None
            self.beta_itm = args.beta_itm if self.use_tim_loss else None
            self.beta_iadds = None
            self.loss_str = ""
            if args.use_clip_loss:
                self.loss_str += "itc{}".format(str(self.beta_itc))
            if args.use_tim_loss:
                self.loss_str += "itm{}".format(str(self.beta_itm))
        self.use_loss_correction = False  # deprecated (config.py:77)
