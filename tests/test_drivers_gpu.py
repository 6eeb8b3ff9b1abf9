"""End-to-end drivers: the conflict-free train.py / predict.py equivalents on the
synthetic dataset, the reference's checkpoint file set (train.py:76-99) and resume (train.py:101-126)."""
import os

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

SMALL = dict(num_channels=[8, 8, 16, 16], num_freqs=201, hidden=32, segment_length=3200, num_layers=2, num_inputs=3,
             kernel_size=3, dropout=0.0, sample_rate=16000, win_length=25, hop_length=10, n_fft=400)


def _config(tmp_path):
    from speech_enhancement_mi_b200 import train
    with open(os.path.join(os.path.dirname(train.__file__), "config.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["TemporalCRN"] = dict(SMALL, precision="fp32")
    cfg["config"]["checkpoint_dir"] = str(tmp_path / "modules")
    cfg["config"]["max_length"] = 30000
    path = tmp_path / "config.yaml"
    path.write_text(yaml.safe_dump(cfg))
    return str(path)


@pytest.mark.parametrize("engine", ["native", "autograd"])
def test_train_checkpoint_resume_predict(tmp_path, engine):
    from speech_enhancement_mi_b200 import predict, train
    cfg = _config(tmp_path)
    argv = ["TemporalCRN", cfg, "--engine", engine, "--steps", "6", "--items", "8", "--epochs", "1"]
    train.main(argv)
    stage = tmp_path / "modules" / "denoise" / "model"
    for name in ("TemporalCRN.pth", "optimizer.pth", "scheduler.pth", "Epoch.pth"):
        assert (stage / name).exists(), name
    sd = torch.load(stage / "TemporalCRN.pth")
    assert "convlist.0.net.0.weight" in sd and len(sd) == 130  # reference key set incl. the alias keys
    opt = torch.load(stage / "optimizer.pth")
    assert len(opt["state"]) >= 102 and {"step", "exp_avg", "exp_avg_sq"} <= set(opt["state"][0])
    assert torch.load(stage / "Epoch.pth")["Epoch"] == 0
    # resume continues from epoch 1 with the saved moments (train.py:252-253)
    train.main(["TemporalCRN", cfg, "--engine", engine, "--steps", "4", "--items", "8", "--epochs", "2", "--resume"])
    assert torch.load(stage / "Epoch.pth")["Epoch"] == 1
    res = predict.predict(type("A", (), dict(name="TemporalCRN", config_path=cfg, user_defined_name="model", items=3))())
    assert np.isfinite(res["si_snr_after_db"]) and res["real_time_factor"] > 0


def test_native_and_autograd_engines_take_the_same_first_step(tmp_path):
    """One optimizer step (2 micro-steps, clip 5, Adam 3e-4) through NativeTrainer equals the reference's literal torch
    loop on the drop-in model (train.py:195-204)."""
    from speech_enhancement_mi_b200 import train
    cfg = _config(tmp_path)
    out = {}
    for engine in ("native", "autograd"):
        args = type("A", (), dict(name="TemporalCRN", config_path=cfg, user_defined_name=engine, engine=engine, steps=2,
                                  items=4, epochs=1, resume=False))()
        p = train.Processor(args)
        p.epoch = -1
        p.run_epoch("train", 2)
        out[engine] = {k: v.detach().cpu().clone() for k, v in p.model.state_dict().items()}
    for k in out["native"]:
        a, b = out["native"][k], out["autograd"][k]
        assert (a - b).abs().max() <= 2e-5, k  # lr 3e-4: one Adam step moves every weight by at most 3e-4


def test_distillation_driver_trains_and_serves_the_student(tmp_path):
    """train_distillation.py:126-215 equivalent (teacher + student + connectors through the native autograd nodes), its
    checkpoint, and the predict_distillation.py:32-38,84 flow: load with strict=False, move to the CPU, run
    ``model.student.realtime_process`` on a CPU mixture in eval mode."""
    from speech_enhancement_mi_b200 import distillation_crn, train_distillation
    from speech_enhancement_mi_b200.data_synth import SyntheticPartyDataset
    cfg = _config(tmp_path)
    train_distillation.main(["DistillationCRN", cfg, "--steps", "4", "--items", "4", "--epochs", "1"])
    stage = tmp_path / "modules" / "distillation" / "model"
    for name in ("DistillationCRN.pth", "optimizer.pth", "scheduler.pth", "Epoch.pth"):
        assert (stage / name).exists(), name
    sd = torch.load(stage / "DistillationCRN.pth")
    assert any(k.startswith("teacher.") for k in sd) and any(k.startswith("student.") for k in sd)
    assert "connectors.0.0.weight" in sd and "connectors.4.1.running_mean" in sd
    with open(cfg) as f:
        full = yaml.safe_load(f)
    kw = full["TemporalCRN"]
    torch.manual_seed(full["config"]["seed"])  # the driver's initialisation (train.py:52)
    before = {k: v.clone() for k, v in distillation_crn.DistillationCRN(**kw).student.state_dict().items()}
    model = distillation_crn.DistillationCRN(**kw).cuda()
    model.load_state_dict(sd, strict=False)
    moved = sum(1 for k, v in model.student.state_dict().items()
                if v.shape == before[k].shape and not torch.equal(v.cpu(), before[k]))
    assert moved > 50  # the optimizer steps reached the student
    model = model.cpu().eval()
    data = SyntheticPartyDataset(size=2, max_length=30000, num_mic=3)
    data.set_attribute("test")
    mix = data[0]["mix"][None]
    with torch.no_grad():
        separated, feats = model.student.realtime_process(mix, flag=False)
    assert feats == [] and separated.shape == (1, mix.shape[-1]) and bool(torch.isfinite(separated).all())
