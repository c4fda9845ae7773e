"""The reference's dataset loaders with the decode / resize work on the device.

    reference (host)                                              here
    utils.get_img        utils.py:17-25   cv2.imread(path, -1)     JPEG: compressed bytes -> nvJPEG on the GPU (ofd_jpeg_decode);
                         -> float32 CHW (B, G, R)                  PNG: host decode (deflate is serial), uint8 over PCIe, widened on the device
    utils.get_depth      utils.py:47-59   imread GRAYSCALE          host PNG decode, the 1 B/px payload goes to the device;
                         .astype(float) + smooth_closer            ofd_depth_from_png does the float64 arithmetic there
    utils.get_disparity  utils.py:61-72   imread UNCHANGED * 63/255  same (8 or 16 bit payload)
    T.Resize(img_size)   dataloader.py:31-32,57-58                  ofd_resize_bilinear_aa (antialiased bilinear, float64)
    ReDWeb / DIML        dataloader.py:14-60                        same directory layout and list files; items are CUDA tensors

`ReDWeb[idx]` -> (img[3,H,W] float32, depth[1,H,W]) and `DIML[idx]` -> (img0, img1, disp0[1,H,W]) feed
preprocess.PreprocessPlusAugment directly: depth / disparity are handed over as the raw uint8 / uint16 payload when no resize is needed (the
driver decodes it with ofd_depth_from_png), or as the float64 tensor the reference's loader would deliver after T.Resize.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _read_bytes(path: str) -> bytes:
    with open(path, "rb") as f:
        return f.read()


class DeviceLoader:
    """One per process and device: owns the nvJPEG decoder."""

    def __init__(self, device=0):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._jpeg = None

    def _decoder(self):
        if self._jpeg is None:
            self._jpeg = ops.JpegDecoder(self.device)
        return self._jpeg

    def get_img(self, path: str):
        """utils.get_img (utils.py:17-25): (img[3,h,w] float32 CUDA in B, G, R order, (h, w))."""
        data = _read_bytes(path)
        if data[:2] == b"\xff\xd8":  # JPEG: decoded on the device from the compressed bytes
            img = self._decoder().decode(data)
            return img, (img.shape[1], img.shape[2])
        import cv2

        arr = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
        if arr is None:
            raise ValueError(f"{path}: not an image cv2 can decode")
        if arr.ndim == 2:
            arr = cv2.cvtColor(arr, cv2.COLOR_GRAY2BGR)
        dev = torch.from_numpy(np.ascontiguousarray(arr[..., :3])).to(self.device, non_blocking=True)  # 1 B per channel over PCIe
        return dev.permute(2, 0, 1).to(torch.float32).contiguous(), arr.shape[:2]

    def _payload(self, path: str, flag):
        import cv2

        arr = cv2.imdecode(np.frombuffer(_read_bytes(path), np.uint8), flag)
        if arr is None:
            raise ValueError(f"{path}: not an image cv2 can decode")
        if arr.ndim == 3:
            arr = arr[..., 0]
        if arr.dtype not in (np.uint8, np.uint16):
            raise TypeError(f"{path}: {arr.dtype} payload (expected 8 or 16 bit)")
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device, non_blocking=True)[None], arr.shape[:2]

    def get_depth_payload(self, path: str):
        """The PNG payload utils.get_depth starts from (utils.py:48): uint8 [1,h,w] CUDA; ops.depth_from_png(payload, "reldepth") is
        utils.get_depth(path, smooth=True)'s float64 result."""
        import cv2

        return self._payload(path, cv2.IMREAD_GRAYSCALE)

    def get_disparity_payload(self, path: str):
        """The PNG payload utils.get_disparity starts from (utils.py:62): uint8 / uint16 [1,h,w] CUDA."""
        import cv2

        return self._payload(path, cv2.IMREAD_UNCHANGED)

    def get_depth(self, path: str, smooth: bool = True):
        """utils.get_depth(path, smooth=True) (utils.py:47-59) as a float64 CUDA tensor [1,h,w]."""
        if not smooth:
            raise NotImplementedError("the reference's drivers always call get_depth with smooth=True")
        payload, size = self.get_depth_payload(path)
        return ops.depth_from_png(payload.contiguous(), "reldepth"), size


class ReDWeb:
    """dataloader.ReDWeb (dataloader.py:14-34): Imgs/<name>.jpg + RDs/<name>.png, names from ReDWeb_list.txt."""

    def __init__(self, dataset_dir="datasets/ReDWeb_V1", list_file="ReDWeb_list.txt", device=0):
        self.dataset_dir = dataset_dir
        with open(list_file, "r") as f:
            self.img_names = [ln.strip() for ln in f if ln.strip()]
        self.loader = DeviceLoader(device)

    def __len__(self):
        return len(self.img_names)

    def __getitem__(self, idx):
        name = self.img_names[idx].split(".")[0]
        img, img_size = self.loader.get_img(f"{self.dataset_dir}/Imgs/{name}.jpg")
        payload, depth_size = self.loader.get_depth_payload(f"{self.dataset_dir}/RDs/{name}.png")
        if tuple(img_size) != tuple(depth_size):  # dataloader.py:31-32: T.Resize(img_size)(depth) on the float64 depth
            depth = ops.resize_bilinear_aa(ops.depth_from_png(payload.contiguous(), "reldepth"), img_size)
            return img, depth
        return img, payload.contiguous()


class DIML:
    """dataloader.DIML (dataloader.py:37-60): train/LR/{outleft,outright,disparity}/<name>.png, names from DIML_list.txt."""

    def __init__(self, dataset_dir="datasets/DIML", list_file="DIML_list.txt", device=0):
        self.dataset_dir = dataset_dir
        with open(list_file, "r") as f:
            self.img_names = [ln.strip() for ln in f if ln.strip()]
        self.loader = DeviceLoader(device)

    def __len__(self):
        return len(self.img_names)

    def __getitem__(self, idx):
        name = self.img_names[idx].split(".")[0]
        base = f"{self.dataset_dir}/train/LR"
        img0, img_size = self.loader.get_img(f"{base}/outleft/{name}.png")
        img1, _ = self.loader.get_img(f"{base}/outright/{name}.png")
        payload, disp_size = self.loader.get_disparity_payload(f"{base}/disparity/{name}.png")
        if tuple(img_size) != tuple(disp_size):  # dataloader.py:57-58: T.Resize on the float64 disparity (disp * 63 / 255, utils.py:66)
            disp = payload.to(torch.float64) * 63 / 255
            return img0, img1, ops.resize_bilinear_aa(disp.contiguous(), img_size)
        return img0, img1, payload.contiguous()
