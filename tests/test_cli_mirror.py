"""CPU: the CLI mirror keeps the reference's flag set (models/run_mm_late.py:20-43) and output file patterns (:87-96,
:117-187).  The `reference`-marked test reads the flags out of the reference's own source (build container only)."""
import ast
import os
import types

import pytest

from tic_b200 import run_mm_late as cli

REF_CLI = "/root/reference/models/run_mm_late.py"


def test_readme_command_parses():
    # README.md:35-38 of the reference
    a = cli.build_parser().parse_args("--txt_model_name bernice --img_model_name vit --fusion_name attention --task 2 "
                                      "--epochs 7 --seed 40 --testing".split())
    assert (a.txt_model_name, a.img_model_name, a.fusion_name, a.task, a.epochs, a.seed, a.testing) == \
        ("bernice", "vit", "attention", 2, 7, 40, True)
    assert a.beta_itc == 0.1 and a.beta_itm == 0.1 and a.itm_rng == "numpy" and not a.use_clip_loss


def test_output_file_patterns():
    a = cli.build_parser().parse_args("--txt_model_name bernice --img_model_name vit --fusion_name concat --task 3 --seed 30 "
                                      "--nsamples 500 --save_model".split())
    cfg = types.SimpleNamespace(loss_str="itc0.1itm0.1")
    model_path, val_csv, te_csv = cli.output_names(a, cfg, "res/")
    stem = "res/bernice-vit-concat_task3_seed30_itc0.1itm0.1_N500_"
    assert (model_path, val_csv, te_csv) == (stem + "net.pth", stem + "metrics_val.csv", stem + "metrics_test.csv")
    names = cli.aux_names(a, cfg, "res/")
    assert names == {k: stem + v for k, v in (("preds", "preds.csv"), ("preds_txt", "preds_txt.csv"),
                                              ("metrics_txt", "metrics_txt.csv"), ("preds_lm", "preds_lm.csv"),
                                              ("metrics_lm", "metrics_lm.csv"))}
    a2 = cli.build_parser().parse_args("--txt_model_name bert --img_model_name vit --fusion_name gmu --task 0".split())
    assert cli.output_names(a2, types.SimpleNamespace(loss_str=""), "r/") == \
        (None, "r/bert-vit-gmu_task0_seed30__metrics_val.csv", "r/bert-vit-gmu_task0_seed30__metrics_test.csv")


def _flags_of(source):
    flags = {}
    for node in ast.walk(ast.parse(source)):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" and node.args:
            name = node.args[0].value
            kw = {k.arg: k.value for k in node.keywords}
            choices = sorted(ast.literal_eval(kw["choices"])) if "choices" in kw else None
            default = ast.literal_eval(kw["default"]) if "default" in kw else None
            action = ast.literal_eval(kw["action"]) if "action" in kw else None
            flags[name] = (choices, default, action)
    return flags


@pytest.mark.reference
def test_flag_set_equals_the_reference():
    ref = _flags_of(open(REF_CLI).read())
    mine = _flags_of(open(cli.__file__).read())
    assert len(ref) >= 20
    extra = set(mine) - set(ref)
    assert extra == {"--itm_rng"} and not (set(ref) - set(mine))
    for name, spec in ref.items():
        assert mine[name] == spec, name
