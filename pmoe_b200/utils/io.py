"""Checkpoint I/O with the reference's file layout (reference: PMoE/utils/io.py:9-45; trainers' `save()`:
train_0.py:313-338 -> {"epoch", "iteration", "unet", ["unet-swa"], "optimizer", "lr_scheduler", "best", ...}, train_1/2.py ->
"model" / "model-swa"). The module classes of this package keep the reference's state_dict keys, FusedAdam / FusedRMSprop
keep torch's optimizer-state names and `optim.AveragedModel` is a `torch.optim.swa_utils.AveragedModel`, so files written
by either side load on the other with strict=True."""
import os
import shutil

import numpy as np
import torch


def save_checkpoint(state: dict, is_best: bool, save_dir: str, name: str):
    """torch.save(state, <save_dir>/<name>.pth); a copy named <prefix>-best.pth when is_best (io.py:9-33)."""
    filepath = os.path.join(save_dir, "%s.pth" % name)
    os.makedirs(save_dir, exist_ok=True)
    torch.save(state, filepath)
    if is_best:
        shutil.copyfile(filepath, os.path.join(save_dir, "%s-best.pth" % name.split("-")[0]))
    return filepath


def load_checkpoint(save: str, device: str):
    """torch.load(save, map_location=device) (io.py:36-47); a missing file is an error, as in the reference."""
    if not os.path.exists(save):
        raise FileNotFoundError("File doesn't exist {}".format(save))
    # the trainers' checkpoints carry numpy scalars next to the state dicts (`best`, `e_loss` = np.mean(...): train_2.py:346-370),
    # which torch >= 2.6 refuses under its weights_only default; these files are the user's own training output
    return torch.load(save, map_location=device, weights_only=False)


def worker_init_fn(worker_id):
    np.random.seed(np.random.get_state()[1][0] + worker_id)
