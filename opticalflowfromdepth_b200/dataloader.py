"""Training-side reader of the files `preprocess.PreprocessPlusAugment` writes (dataloader.py:60-232 of
the reference): `AugmentedDataset.getitem_from_npz` and `DepthToFlowDataset.getitem_from_npz` with the reference's
channel slicing, normalisation, flips, crop and one-hot label — minus its defects:

  * the reference reads a key `augment_img` its writer never stores and then recurses forever through a bare `except`;
    here the key is optional — absent, it is inferred from the file name (`*_1.npz` holds the augmented FIRST image, `*_2.npz`
    the augmented second one, preprocess.py:459-476) — and real I/O errors propagate;
  * `DepthToFlowDataset` crops with undefined `h, w` (dataloader.py:221); here they come from the group tensor.

Pure host code (numpy / torch CPU): the arrays come off disk; nothing here is on the GPU hot path.  Random decisions use
`np.random` in the reference's order (h-flip, v-flip, crop y0, crop x0), so a seeded run reproduces the reference's sample.
"""
from __future__ import annotations

import re

import numpy as np
import torch
from torch.utils import data

num_classes = 1 + 3  # dataloader.py:11


def _to_tensor(chw: np.ndarray, size=None) -> torch.Tensor:
    """The reader's transform (dataloader.py:68-77,104-106,118): T.ToTensor() on the float HWC view of a CHW array - a float array
    is not rescaled, so that part is the identity on CHW data - followed by T.Resize(size) when `size` is given (torchvision's
    host-side resize with its default arguments, exactly the reference's call)."""
    t = torch.from_numpy(np.ascontiguousarray(chw))
    if size is not None:
        import torchvision.transforms as T

        t = T.Resize(size)(t)
    return t


def _flip_and_crop(img0, img1, img0_depth, flow, h, w, do_flip, h_flip_prob, v_flip_prob, crop_size):
    if do_flip:
        if np.random.rand() < h_flip_prob:  # dataloader.py:130-135
            img0, img1, img0_depth, flow = (torch.flip(t, (2,)) for t in (img0, img1, img0_depth, flow))
            flow[0] = flow[0] * -1.0
        if np.random.rand() < v_flip_prob:  # :137-142
            img0, img1, img0_depth, flow = (torch.flip(t, (1,)) for t in (img0, img1, img0_depth, flow))
            flow[1] = flow[1] * -1.0
    if crop_size is not None:  # :144-151
        y0 = np.random.randint(0, h - crop_size[0] + 1)
        x0 = np.random.randint(0, w - crop_size[1] + 1)
        sl = (slice(None), slice(y0, y0 + crop_size[0]), slice(x0, x0 + crop_size[1]))
        img0, img1, img0_depth, flow = img0[sl], img1[sl], img0_depth[sl], flow[sl]
    return img0, img1, img0_depth, flow


def _one_hot(label_type: int) -> torch.Tensor:
    label = torch.zeros(num_classes)
    label[label_type] = 1
    return label


class AugmentedDataset(data.Dataset):
    """dataloader.AugmentedDataset (dataloader.py:60-157).  `size`: every tensor goes through T.Resize(size) after the channel slicing, as
    in the reference (flips and crop offsets still use the file's h and w, dataloader.py:84,144-146)."""

    def __init__(self, normalize_dataset=True, size=None, crop_size=None, do_flip=True):
        self.size = size
        self.normalize_dataset = normalize_dataset
        self.crop_size = crop_size
        self.do_flip = do_flip
        self.h_flip_prob = 0.5
        self.v_flip_prob = 0.1

    def getitem_from_npz(self, npz_filename, group_npz_filename, random_group, idx=None):
        npz_file = np.load(npz_filename)
        if "augment_img" in npz_file.files:
            augment_img = int(npz_file["augment_img"])
        else:
            m = re.search(r"_([12])\.npz$", str(npz_filename))
            if not m:
                raise KeyError(f"{npz_filename}: no `augment_img` key and the name does not end in _1.npz / _2.npz")
            augment_img = int(m.group(1)) - 1
        augment_flow_type = int(npz_file["augment_flow_type"])
        img_depth_flow = np.array(npz_file["img_depth_flow"])  # a private copy: normalised in place below
        _, h, w = img_depth_flow.shape
        group = np.load(group_npz_filename)["img_depth_flow"]
        if random_group == 0:    # dataloader.py:93-104
            img0, img0_depth, img1 = group[0:3], group[3:4], group[4:7]
        elif random_group == 1:
            img0, img0_depth, img1 = group[4:7], group[7:8], group[8:11]
        elif random_group == 2:
            img0, img0_depth, img1 = group[0:3], group[3:4], group[8:11]
        else:
            raise ValueError("random_group must be 0, 1 or 2")
        img0, img1, img0_depth = _to_tensor(img0, self.size), _to_tensor(img1, self.size), _to_tensor(img0_depth, self.size)
        if self.normalize_dataset:  # :108-116 (flow.x / h and flow.y / w, as the reference has it)
            if augment_img == 0:
                img_depth_flow[4] = img_depth_flow[4] / h
                img_depth_flow[5] = img_depth_flow[5] / w
                img_depth_flow[3] = img_depth_flow[3] / 100
            else:
                img_depth_flow[0] = img_depth_flow[0] / h
                img_depth_flow[1] = img_depth_flow[1] / w
                img_depth_flow[7] = img_depth_flow[7] / 100
        t = _to_tensor(img_depth_flow, self.size)
        if augment_img == 0:  # :120-126
            img0, img0_depth, flow = t[0:3], t[3:4], t[4:6]
        else:
            flow, img1 = t[0:2], t[4:7]
        img0, img1, img0_depth, flow = _flip_and_crop(img0, img1, img0_depth, flow, h, w, self.do_flip, self.h_flip_prob,
                                                      self.v_flip_prob, self.crop_size)
        return img0, img1, flow, img0_depth, _one_hot(max(0, augment_flow_type - 4))


class DepthToFlowDataset(data.Dataset):
    """dataloader.DepthToFlowDataset (dataloader.py:160-232): pairs straight from group.npz, label 0."""

    def __init__(self, normalize_dataset=True, size=None, crop_size=None, do_flip=True):
        self.size = size
        self.normalize_dataset = normalize_dataset
        self.crop_size = crop_size
        self.do_flip = do_flip
        self.h_flip_prob = 0.5
        self.v_flip_prob = 0.1

    def getitem_from_npz(self, group_npz_filename, random_group, idx=None):
        group = np.load(group_npz_filename)["img_depth_flow"]
        _, h, w = group.shape
        if random_group == 0:    # dataloader.py:185-200
            img0, img0_depth, img1, flow = group[0:3], group[3:4], group[4:7], group[12:14]
        elif random_group == 1:
            img0, img0_depth, img1, flow = group[4:7], group[7:8], group[8:11], group[16:18]
        elif random_group == 2:
            img0, img0_depth, img1, flow = group[0:3], group[3:4], group[8:11], group[20:22]
        else:
            raise ValueError("random_group must be 0, 1 or 2")
        img0, img1, img0_depth, flow = (_to_tensor(np.array(a), self.size) for a in (img0, img1, img0_depth, flow))
        img0, img1, img0_depth, flow = _flip_and_crop(img0, img1, img0_depth, flow, h, w, self.do_flip, self.h_flip_prob,
                                                      self.v_flip_prob, self.crop_size)
        return img0, img1, flow, img0_depth, _one_hot(0)


class AugmentedFolder(AugmentedDataset):
    """AugmentedDIML / AugmentedReDWeb (dataloader.py:235-268) over any output directory of the preprocess driver:
    item idx draws a random pair group (0..2), augmentation (0..11) and set (1..2) from `{root}/{idx}/`."""

    def __init__(self, root, n_frames, normalize_dataset=True, size=None, crop_size=None):
        super().__init__(normalize_dataset=normalize_dataset, size=size, crop_size=crop_size)
        self.root, self.n_frames = root, n_frames

    def __len__(self):
        return self.n_frames

    def __getitem__(self, idx):
        random_group = np.random.randint(0, 3)
        random_augment = np.random.randint(0, 12)
        random_set = np.random.randint(1, 3)
        d = f"{self.root}/{idx}"
        return self.getitem_from_npz(f"{d}/{random_group}_{random_augment}_{random_set}.npz", f"{d}/group.npz", random_group, idx)
