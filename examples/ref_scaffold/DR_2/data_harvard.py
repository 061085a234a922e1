"""Synthetic stand-in for ``DR_2.data_harvard.GAMMA_dataset`` (unpublished; the published sibling is
code/data_harvard.py:598-857, whose constructor and output contract this follows).

No dataset is available offline, so every item is generated: a raw fundus photograph ``uint8 [200, 200, 3]`` and a raw
OCT volume ``uint8 [200, 200, 200]`` (BASELINE.json: "Harvard-30K-shaped"), then the reference's own preprocessing --
``cv2.resize`` to 384 x 384 INTER_CUBIC, nearest-neighbour zoom to 96^3, ``/ 255`` (code/data_harvard.py:169-183,686-695)
-- and its two views under ``--condition noise --condition_name Gaussian``: ``np.random.seed(seed_idx)`` per item, a
zero-variance draw for the clean view, ``clip(x + N(0, 0.5), 0, 1)`` for the noisy one, OCT before fundus
(code/data_harvard.py:698,722-731,769-783), ``ToTensor`` (+ the train-mode flips / colour jitter) and
``((low, high), label)`` with ``low = {0: fundus [3,384,384], 1: oct [1,96,96,96]}`` (code/data_harvard.py:817-841).

``EDRL_SYNTH_POOL=P`` (default 4): each worker keeps P preprocessed base items and cycles through them, so that the
loader measures the reference's per-item view generation and tensor plumbing, not ``numpy`` drawing 8 M raw voxels;
``0`` generates every item from scratch.  ``EDRL_SYNTH_MISSING=oct|fundus`` zero-fills that modality (BASELINE
configs[4]: missing-modality inference; the reference has no such switch, zero filling is its commented-out path
code/data_harvard.py:280,334).
"""
import os

import cv2
import numpy as np
import torch
from scipy import ndimage
from torch.utils.data import Dataset
from torchvision import transforms


def scale_image(image, patch_size):
    return cv2.resize(image, (patch_size, patch_size), interpolation=cv2.INTER_CUBIC)


def resize_oct_data_trans(data, size):
    depth, height, width = data.shape
    scale = [size[0] * 1.0 / depth, size[1] * 1.0 / height, size[2] * 1.0 / width]
    return ndimage.zoom(data, scale, order=0)


class GAMMA_dataset(Dataset):
    def __init__(self, args, dataset_root, oct_img_size, fundus_img_size, mode='train', label_file='', filelists=None):
        self.condition = args.condition
        self.condition_name = args.condition_name
        self.seed_idx = args.seed_idx
        self.model_base = args.model_base
        self.mode = mode.lower()
        self.missing = os.environ.get("EDRL_SYNTH_MISSING", "")
        self.pool_size = int(os.environ.get("EDRL_SYNTH_POOL", "4"))
        self._pool = {}
        self.fundus_train_transforms = transforms.Compose([
            transforms.ToTensor(),
            transforms.RandomApply([transforms.ColorJitter(0.2, 0.2, 0.2, 0.1)], p=0.8),
            transforms.RandomGrayscale(p=0.2),
            transforms.RandomHorizontalFlip(),
        ])
        self.oct_train_transforms = transforms.Compose([transforms.ToTensor(), transforms.RandomHorizontalFlip()])
        self.val_transforms = transforms.Compose([transforms.ToTensor()])
        # labels: the reference reads one-hot rows from an xlsx keyed by the numeric file name; here the name's parity
        self.file_list = []
        for f in filelists:
            name = os.path.basename(str(f))
            if name.isdigit():
                onehot = np.zeros(2, dtype=np.int64)
                onehot[int(name) % 2] = 1
                self.file_list.append([name, onehot])

    def __len__(self):
        return len(self.file_list)

    def _base_item(self, key):
        rng = np.random.default_rng(int(key))
        fundus_raw = rng.integers(0, 256, size=(200, 200, 3), dtype=np.uint8)
        oct_raw = rng.integers(0, 256, size=(200, 200, 200), dtype=np.uint8).astype(np.float32)
        if self.model_base == "transformer":
            fundus = scale_image(fundus_raw, 384)
            oct_img = resize_oct_data_trans(oct_raw, (96, 96, 96))
        else:
            fundus = scale_image(fundus_raw, 512)
            oct_img = resize_oct_data_trans(oct_raw, (128, 256, 128))
        return fundus / 255.0, oct_img / 255.0

    def __getitem__(self, idx):
        real_index, label = self.file_list[idx]
        key = int(real_index) % self.pool_size if self.pool_size > 0 else int(real_index)
        if key not in self._pool:
            item = self._base_item(key)
            if self.pool_size > 0:
                self._pool[key] = item
        else:
            item = self._pool[key]
        fundus_img, oct_img = item
        if self.missing == "oct":
            oct_img = np.zeros_like(oct_img)
        elif self.missing == "fundus":
            fundus_img = np.zeros_like(fundus_img)
        np.random.seed(self.seed_idx)
        if self.condition == 'noise':
            oct_low = np.clip(oct_img + np.random.normal(0, 0, oct_img.shape), 0.0, 1.0)
            fundus_low = np.clip(fundus_img + np.random.normal(0, 0, fundus_img.shape), 0.0, 1.0)
            oct_high = np.clip(oct_img + np.random.normal(0, 0.5, oct_img.shape), 0.0, 1.0)
            fundus_high = np.clip(fundus_img + np.random.normal(0, 0.5, fundus_img.shape), 0.0, 1.0)
        else:
            fundus_low = fundus_high = fundus_img
            oct_low = oct_high = oct_img
        if self.mode == "train":
            f_lo = self.fundus_train_transforms(fundus_low.astype(np.float32))
            o_lo = self.oct_train_transforms(oct_low.astype(np.float32))
            f_hi = self.fundus_train_transforms(fundus_high.astype(np.float32))
            o_hi = self.oct_train_transforms(oct_high.astype(np.float32))
        else:
            f_lo, o_lo = self.val_transforms(fundus_low), self.val_transforms(oct_low)
            f_hi, o_hi = self.val_transforms(fundus_high), self.val_transforms(oct_high)
        data_low = {0: f_lo, 1: o_lo.unsqueeze(0)}
        data_high = {0: f_hi, 1: o_hi.unsqueeze(0)}
        return (data_low, data_high), label.argmax()
