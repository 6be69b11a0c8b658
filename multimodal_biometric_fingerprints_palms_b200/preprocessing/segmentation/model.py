"""Mirror of `src/preprocessing/segmentation/model.py` of the reference (model.py:8-99) for INFERENCE: the same class
names and state_dict layout, the forward pass computed by libfpb200 (include/fpb200_unet.h: tcgen05 implicit-GEMM
convolutions, eval-mode BatchNorm folded in).  Training (src/preprocessing/segmentation/train.py) is out of scope."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional

import numpy as np

from ... import _native as N

# ConvBlocks of NestedUNet.__init__ (model.py:34-58) in the order of the FPB_UNET_* enum
BLOCKS = ("conv0_0", "conv1_0", "conv2_0", "conv3_0", "conv4_0", "up1_0", "up2_0", "up3_0", "up1_1", "up2_1", "up1_2")


def _np32(v) -> np.ndarray:
    if hasattr(v, "detach"):                     # torch.Tensor
        v = v.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(v, dtype=np.float32))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class NestedUNet:
    """model.py:26-83.  `NestedUNet(num_labels=1, input_channels=3)`, then `load_state_dict(...)` with the reference's
    keys (`conv0_0.conv.0.weight` ... `final.bias`), then `model(x)` with x float32 [N, C, H, W] (NumPy or torch, values as
    the reference feeds them: grey / 255 replicated to three channels, inference.py:87-93) -> logits [N, 1, H, W].
    The CUDA engine is created for the first input's (H, W) (multiples of 16)."""

    def __init__(self, num_labels: int = 1, input_channels: int = 3, deep_supervision: bool = False, device: int = 0,
                 max_batch: int = 8):
        if num_labels != 1:
            raise NotImplementedError("CUDA path implements num_labels=1 (config_segmentation.yml model.num_labels)")
        self.num_labels, self.input_channels, self.deep_supervision = num_labels, input_channels, deep_supervision
        self.device, self.max_batch = int(device), int(max_batch)
        self._lib = N.load()
        self._h = C.c_void_p()
        self._shape = None
        self._params: Optional[Dict[str, np.ndarray]] = None
        self.training = False

    # ---- torch.nn.Module look-alikes the reference's inference script calls (inference.py:80-83)
    def eval(self):
        self.training = False
        return self

    def to(self, *_a, **_k):
        return self

    def load_state_dict(self, state: Mapping, strict: bool = True):
        sd = {k[len("model."):] if k.startswith("model.") else k: _np32(v) for k, v in state.items()
              if not k.endswith("num_batches_tracked")}
        want = []
        for b in BLOCKS:
            for i in (0, 1, 3, 4):
                want += [f"{b}.conv.{i}.weight", f"{b}.conv.{i}.bias"]
            for i in (1, 4):
                want += [f"{b}.conv.{i}.running_mean", f"{b}.conv.{i}.running_var"]
        want += ["final.weight", "final.bias"]
        missing = [k for k in want if k not in sd]
        if missing and strict:
            raise RuntimeError(f"Missing key(s) in state_dict: {missing[:6]}{' ...' if len(missing) > 6 else ''}")
        self._params = sd
        if self._h:
            self._upload()

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise N.FpbError(f"{what} failed ({rc}): {self._lib.fpb_unet_last_error(self._h).decode()}")

    def _create(self, H: int, W: int):
        self.close()
        rc = self._lib.fpb_unet_create(C.byref(self._h), self.device, self.max_batch, H, W, self.input_channels)
        if rc != 0:
            raise N.FpbError(f"fpb_unet_create failed ({rc}): {self._lib.fpb_unet_last_error(None).decode()}")
        self._shape = (H, W)
        if self._params is not None:
            self._upload()

    def _upload(self):
        sd = self._params
        for bi, b in enumerate(BLOCKS):
            for ci, (conv, bn) in enumerate(((0, 1), (3, 4))):
                ic, oc = C.c_int(), C.c_int()
                self._ck(self._lib.fpb_unet_conv_shape(self._h, bi, ci, C.byref(ic), C.byref(oc)), "fpb_unet_conv_shape")
                w = sd[f"{b}.conv.{conv}.weight"]
                if w.shape != (oc.value, ic.value, 3, 3):
                    raise RuntimeError(f"size mismatch for {b}.conv.{conv}.weight: {w.shape} vs {(oc.value, ic.value, 3, 3)}")
                arrs = [w, sd[f"{b}.conv.{conv}.bias"], sd[f"{b}.conv.{bn}.weight"], sd[f"{b}.conv.{bn}.bias"],
                        sd[f"{b}.conv.{bn}.running_mean"], sd[f"{b}.conv.{bn}.running_var"]]
                self._ck(self._lib.fpb_unet_set_conv(self._h, bi, ci, *[_p(a) for a in arrs], 1e-5), "fpb_unet_set_conv")
        fw, fb = sd["final.weight"].reshape(-1), sd["final.bias"].reshape(-1)
        if fw.size != 64 or fb.size != 1:
            raise RuntimeError("final convolution must be Conv2d(64, 1, 1)")
        self._ck(self._lib.fpb_unet_set_final(self._h, _p(np.ascontiguousarray(fw)), _p(np.ascontiguousarray(fb))), "fpb_unet_set_final")

    def forward(self, x):
        is_torch = hasattr(x, "detach")
        a = _np32(x)
        if a.ndim != 4 or a.shape[1] != self.input_channels:
            raise ValueError(f"expected [N, {self.input_channels}, H, W], got {a.shape}")
        if self._params is None:
            raise RuntimeError("load_state_dict first: this inference engine has no parameter initialisation of its own")
        n, _, H, W = a.shape
        if self._shape != (H, W):
            self._create(H, W)
        out = np.empty((n, 1, H, W), np.float32)
        for s in range(0, n, self.max_batch):
            e = min(n, s + self.max_batch)
            chunk = np.ascontiguousarray(a[s:e])
            self._ck(self._lib.fpb_unet_forward(self._h, _p(chunk), e - s, _p(out[s:e])), "fpb_unet_forward")
        if is_torch:
            import torch
            return torch.from_numpy(out)
        return out

    __call__ = forward

    def launches(self):
        t, tc = C.c_int(), C.c_int()
        self._lib.fpb_unet_launches(self._h, C.byref(t), C.byref(tc))
        return t.value, tc.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.fpb_unet_destroy(self._h)
            self._h = C.c_void_p()
            self._shape = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FingerprintSegmentationModel:
    """model.py:89-99: wrapper with `.model = NestedUNet(...)`; `pretrained_model` is ignored there too."""

    def __init__(self, num_labels: int = 1, image_size=(224, 224), pretrained_model=None, device: int = 0, max_batch: int = 8):
        self.image_size = image_size
        self.model = NestedUNet(num_labels=num_labels, input_channels=3, device=device, max_batch=max_batch)

    def load_state_dict(self, state: Mapping, strict: bool = True):
        return self.model.load_state_dict(state, strict)

    def eval(self):
        self.model.eval()
        return self

    def to(self, *_a, **_k):
        return self

    def forward(self, x):
        return self.model(x)

    __call__ = forward
