"""Conflict-free equivalent of the reference's ``predict.py`` (merge markers at predict.py:4-9,18-27) for CRN_ELU:
load ``<checkpoint_dir>/denoise/<name>/TemporalCRN.pth``, enhance test mixtures with ``realtime_process`` and report
SI-SNR before / after plus the real-time factor (predict.py:45-48,71-72,92).

    python -m speech_enhancement_mi_b200.predict TemporalCRN config.yaml [--user_defined_name model] [--items 8]
"""
from __future__ import annotations

import argparse
import os
import time

import torch
import yaml

from . import CRN_ELU
from .data_synth import SyntheticPartyDataset
from .utility import cal_si_snr


def predict(args):
    with open(args.config_path, "r", encoding="utf-8") as f:
        config = yaml.load(f.read(), Loader=yaml.FullLoader)
    stage_dir = os.path.join(config["config"]["checkpoint_dir"], "denoise", args.user_defined_name)
    model = getattr(CRN_ELU, args.name)(**config[args.name])
    path = os.path.join(stage_dir, args.name + ".pth")
    if not os.path.exists(path):
        if not getattr(args, "allow_random_init", False):
            raise FileNotFoundError(f"{path}: no checkpoint to evaluate (train first, or pass --allow_random_init to "
                                    "measure the real-time factor on random weights)")
        print(f"WARNING: {path} is missing; evaluating RANDOM-INIT weights (quality numbers are meaningless)")
    else:
        model.load_state_dict(torch.load(path), strict=False)  # predict.py:47
    model.eval()
    data = SyntheticPartyDataset(size=args.items, max_length=config["config"]["max_length"])
    data.init_seed(12345)
    before = after = audio = wall = 0.0
    for index in range(len(data)):
        item = data[index]
        mixture, source, length = item["mix"][None], item["source"][None].squeeze(1)[:, 0], item["length"].reshape(-1)
        with torch.no_grad():
            t0 = time.time()
            separated = model.realtime_process(mixture, item["flag"])  # host tensors: copies inside (predict.py:48,92)
            wall += time.time() - t0
        audio += mixture.shape[-1] / 16000.0
        before += float(cal_si_snr(mixture[:, 0], source, length))
        after += float(cal_si_snr(separated, source, length))
    n = len(data)
    res = {"si_snr_before_db": before / n, "si_snr_after_db": after / n, "audio_s": audio, "wall_s": wall,
           "real_time_factor": wall / audio}
    print(res)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("name")
    ap.add_argument("config_path")
    ap.add_argument("--user_defined_name", default="model")
    ap.add_argument("--items", type=int, default=8)
    ap.add_argument("--allow_random_init", action="store_true")
    predict(ap.parse_args())
