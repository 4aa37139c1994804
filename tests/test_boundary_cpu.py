"""Host-side checks of the drop-in Python boundary that need no GPU: config merging, the
factories' error behaviour, state-dict key names, the built-in trainer's hook order, loud failure on CPU."""
import argparse
import os

import pytest
import torch
import torch.nn as nn

import multimodal_clinical_b200 as pkg
from multimodal_clinical_b200 import _lib
from multimodal_clinical_b200.utils import lightning_compat as lc
from multimodal_clinical_b200.utils.merge_yaml import deep_merge, load_and_merge_yaml
from multimodal_clinical_b200.utils.setup_configs import setup_configs

PKG = os.path.dirname(os.path.abspath(pkg.__file__))


def test_deep_merge_overrides_and_recurses():
    a = {"x": 1, "n": {"p": 1, "q": 2}, "keep": 3}
    deep_merge(a, {"x": 5, "n": {"q": 7, "r": 8}})
    assert a == {"x": 5, "n": {"p": 1, "q": 7, "r": 8}, "keep": 3}


def test_setup_configs_merges_dataset_yaml_over_base():
    args = setup_configs(["--dir", "cremad"])
    assert args.num_classes == 6 and args.batch_size == 64 and args.alpha == 0.8 and args.seed == 5
    assert args.grad_mod_type == "OGM_GE" and args.use_scheduler is True
    base = load_and_merge_yaml(os.path.join(PKG, "utils", "base_cfg.yaml"), os.path.join(PKG, "food101", "food101.yaml"))
    assert base["num_classes"] == 101 and base["model_type"] == "qmf" and base["dropout_p"] == 0.1
    with pytest.raises(NotImplementedError):
        setup_configs([])


def test_bad_dir_and_model_type_raise_like_the_reference():
    from multimodal_clinical_b200 import main
    with pytest.raises(NotImplementedError):
        main.main(["--dir", "nope"])
    from multimodal_clinical_b200 import cremad, enrico, food101
    for mod in (cremad, enrico, food101):
        with pytest.raises(NotImplementedError):
            mod.get_model(argparse.Namespace(model_type="not_a_model"))


def _args(**kw):
    d = dict(num_classes=6, num_samples=50, learning_rate=1e-3, use_scheduler=True, grad_mod_type="OGM", alpha=0.8,
             encoder="precomputed")
    d.update(kw)
    return argparse.Namespace(**d)


def test_state_dict_keys_keep_the_reference_names():
    from multimodal_clinical_b200.cremad.joint_model_qmf import MultimodalCremadModel as CQ
    from multimodal_clinical_b200.enrico.joint_model import MultimodalEnricoModel as EJ
    from multimodal_clinical_b200.food101.joint_model_qmf import MultimodalFoodModel as FQ
    keys = set(CQ(_args()).state_dict())
    for k in ("model.x1_classifier.weight", "model.x1_classifier.bias", "model.x2_classifier.weight",
              "model.x1_model.conv1.weight", "model.x2_model.layer4.1.bn2.running_var"):
        assert k in keys, k
    assert not any("fused" in k for k in keys)          # the fused head adds no parameters / buffers
    keys = set(EJ(_args(num_classes=20)).state_dict())
    assert "model.x1_model.classifier.weight" in keys and "model.x2_model.classifier.bias" in keys
    keys = set(FQ(_args(num_classes=101)).state_dict())
    for k in ("model.x1_model.mlp.0.weight", "model.x1_model.mlp.3.bias", "model.x2_model.mlp.6.weight"):
        assert k in keys, k


def test_multi_modality_modules_keep_the_reference_names_and_refuse_cpu():
    """mustard (three modalities) / avmnist (unequal widths): parameter names and shapes of the reference's FusionNets
    (mustard/joint_model.py:9-58, avmnist/joint_model.py:32-110), Adam optimizer, loud failure without a GPU."""
    from multimodal_clinical_b200.mustard.joint_model import MultimodalMustardModel as MM
    from multimodal_clinical_b200.avmnist.joint_model import MultimodalAVMnistModel as AM
    from multimodal_clinical_b200 import mustard, avmnist
    sd = MM(argparse.Namespace(num_classes=2, learning_rate=5e-4)).state_dict()
    assert len(sd) == 30 and not any("fused" in k for k in sd)
    assert tuple(sd["model.x1_model.fc1.weight"].shape) == (384, 371) and tuple(sd["model.x2_model.fc1.weight"].shape) == (384, 81)
    assert tuple(sd["model.x3_model.fc3.weight"].shape) == (2, 100) and "model.x2_model.lstm.weight_hh_l0" in sd
    am = AM(argparse.Namespace(num_classes=10, learning_rate=1e-3))
    sd = am.state_dict()
    assert len(sd) == 64 and tuple(sd["model.classifier_x1.weight"].shape) == (10, 48) and tuple(sd["model.classifier_x2.weight"].shape) == (10, 192)
    assert tuple(sd["model.x2_model.convs.5.weight"].shape) == (192, 96, 3, 3) and "model.x1_model.bns.3.running_mean" in sd
    assert isinstance(am.configure_optimizers(), torch.optim.Adam)
    with pytest.raises(_lib.LfError):
        am.model(torch.randn(4, 1, 28, 28), torch.randn(4, 1, 112, 112), torch.zeros(4, dtype=torch.long))
    for mod in (mustard, avmnist):
        with pytest.raises(NotImplementedError):
            mod.get_model(argparse.Namespace(model_type="jprobas"))


def test_base_model_public_attributes_and_optimizer():
    from multimodal_clinical_b200.cremad.joint_model_ogm_ge import MultimodalCremadModel
    m = MultimodalCremadModel(_args())
    assert m.automatic_optimization is False and m.ogm_modulation == "OGM" and m.ogm_alpha == 0.8
    assert m.num_modality == 2 and tuple(m.ema_offset.x.shape) == (2, 6) and m.ema_offset.smoothing == 0.05
    assert set(m.train_metrics) >= {"train_loss", "train_acc", "train_x1_acc_uncal", "train_x2_acc"}
    opts, scheds = m.configure_optimizers()
    assert isinstance(opts[0], torch.optim.SGD) and opts[0].defaults["momentum"] == 0.9
    assert opts[0].defaults["weight_decay"] == 1e-4 and scheds[0]["scheduler"].step_size == 70


def test_fused_head_refuses_cpu_tensors():
    from multimodal_clinical_b200.heads import FusedLateFusionHead
    h = FusedLateFusionHead(6)
    with pytest.raises(_lib.LfError):
        h(torch.randn(4, 8), torch.randn(4, 8), nn.Linear(8, 6), nn.Linear(8, 6), torch.zeros(4, dtype=torch.long))
    from multimodal_clinical_b200.existing_algos.QMF import QMF
    with pytest.raises(_lib.LfError):
        QMF(2, 10).df(torch.randn(2, 4, 6))


@pytest.mark.skipif(lc.HAVE_LIGHTNING, reason="built-in trainer only used without pytorch_lightning")
def test_builtin_trainer_calls_hooks_in_lightning_order(tmp_path):
    calls = []

    class M(lc.LightningModule):
        def __init__(self):
            super().__init__()
            self.w = nn.Parameter(torch.zeros(1))

        def training_step(self, batch, i):
            calls.append(("train", i)); return (self.w - batch[0].mean()) ** 2

        def validation_step(self, batch, i):
            calls.append(("val", i)); assert not torch.is_grad_enabled()

        def on_validation_epoch_end(self):
            calls.append("val_end"); self.log("val_epoch/val_avg_acc", torch.tensor(float(len(calls))))

        def on_train_epoch_end(self):
            calls.append("train_end")

        def test_step(self, batch, i):
            calls.append(("test", i))

        def on_test_epoch_end(self):
            calls.append("test_end")

        def configure_optimizers(self):
            opt = torch.optim.SGD(self.parameters(), lr=0.1)
            return [opt], [{"scheduler": torch.optim.lr_scheduler.StepLR(opt, 1, 0.5), "interval": "epoch"}]

    data = [(torch.ones(2),), (torch.ones(2) * 3,)]
    ck = lc.ModelCheckpoint(dirpath=str(tmp_path), filename="best", monitor="val_epoch/val_avg_acc", mode="max")
    tr = lc.Trainer(max_epochs=2, callbacks=[ck])
    m = M()
    tr.fit(m, train_dataloaders=data, val_dataloaders=data[:1])
    assert calls[:5] == [("train", 0), ("train", 1), ("val", 0), "val_end", "train_end"]
    assert float(m.w) != 0.0 and abs(tr.optimizers[0].param_groups[0]["lr"] - 0.025) < 1e-12
    assert os.path.exists(ck.best_model_path)
    tr.test(m, dataloaders=data)
    assert calls[-3:] == [("test", 0), ("test", 1), "test_end"]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores; the only bench arm that runs without a GPU):
    one JSON line with the contract's keys, `impl: reference`, zero H2D/D2H bytes and a cpu_baseline describing itself."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "k2",
                          "--steps", "2", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "fusion_step_train_samples_per_sec"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and line["config"]["workload"].startswith("k2")


def test_bench_algorithmic_bytes_follow_survey_8d():
    """roofline.achieved / step_roofline rest on these figures: SURVEY.md 8(d) per-sample bytes (fp32), the two-pass QMF step
    adding one more read of the features, bf16 halving the feature / dfeat terms."""
    import bench
    W = bench.WORKLOADS
    assert bench.step_alg_bytes_per_sample(W["k3"]) == 8272                       # 4096 + 4096 + 72 + 8
    assert bench.step_alg_bytes_per_sample(W["k5"]) == 11908                      # 4096 + 4096 + 3708 + 8
    one_pass_k4 = 6144 + 6144 + 1616 + 16
    assert one_pass_k4 == 13920
    assert bench.step_alg_bytes_per_sample(W["k4"]) == one_pass_k4 + 2 * 768 * 4  # two-pass: + M D 4 = 20 064
    assert bench.step_alg_bytes_per_sample(W["k4"], fe=2) == 10848                # the figure of the default (bf16) bench line
    assert bench.step_alg_bytes_per_sample(W["k2"]) == 8304 + 2 * 512 * 4
    # per-launch bytes of the K4 bf16 kernels: every input read once, every output written once
    B, D, C = W["k4"]["B"], W["k4"]["D"], W["k4"]["C"]
    fwd = bench.kernel_alg_bytes("tc_forward_qmf", W["k4"], B, fe=2)
    assert fwd == (2 * B * D + 2 * C * D) * 2 + (4 * B * C + 6 * B) * 4 + 8 * B == 154975232
    dw = bench.kernel_alg_bytes("tc_dweight", W["k4"], B, fe=2)
    assert dw == (2 * B * 104 + 2 * B * D) * 2 + 2 * C * D * 4 == 114915328
    assert bench.kernel_alg_bytes("no_such_kernel", W["k4"], B) is None
