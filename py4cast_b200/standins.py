"""
Stand-in convolutional forecaster for BASELINE configs #1 / #5.

The reference's CNN / ViT models (HalfUNet, UNet, Segformer ...) live in the absent third-party
package mfai and are OUT OF SCOPE as kernels: those configs only push a
`features_second` model through OUR AR-step and loss kernels.  This small stock-PyTorch
(cuDNN) half-U-Net honours the same plugin contract (`ModelABC`, keyword constructor,
`(B,F,H,W) -> (B,F_out,H,W)`), so `model_name: HalfUNet` configs run when mfai is not installed.
When mfai IS importable its own models take precedence (py4cast_b200/models.py).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch
from torch import nn

from .base import ModelABC, ModelType
from .nlam import _json_dataclass


@_json_dataclass
@dataclass(slots=True)
class HalfUNetSettings:
    num_filters: int = 16
    dilation: int = 1
    bias: bool = False


class HalfUNet(ModelABC, nn.Module):
    settings_kls = HalfUNetSettings
    onnx_supported: bool = True
    supported_num_spatial_dims = (2,)
    num_spatial_dims: int = 2
    features_last: bool = False
    model_type = ModelType.CONVOLUTIONAL
    register: bool = True
    is_stand_in: bool = True

    def __init__(self, in_channels: int, out_channels: int, input_shape: Tuple[int, int],
                 settings: HalfUNetSettings = None, *args, **kwargs):
        super().__init__()
        settings = settings if settings is not None else HalfUNetSettings()
        self._settings = settings
        self.in_channels, self.out_channels, self.input_shape = in_channels, out_channels, tuple(input_shape)
        c = settings.num_filters

        def block(i, o):
            return nn.Sequential(
                nn.Conv2d(i, o, 3, padding=settings.dilation, dilation=settings.dilation, bias=settings.bias),
                nn.BatchNorm2d(o), nn.ReLU(),
                nn.Conv2d(o, o, 3, padding=1, bias=settings.bias), nn.BatchNorm2d(o), nn.ReLU(),
            )

        self.enc1, self.enc2, self.enc3 = block(in_channels, c), block(c, c), block(c, c)
        self.pool = nn.MaxPool2d(2)
        self.up2 = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False)
        self.up4 = nn.Upsample(scale_factor=4, mode="bilinear", align_corners=False)
        self.dec = block(c, c)
        self.head = nn.Conv2d(c, out_channels, 1)
        self.check_required_attributes()

    @property
    def settings(self) -> HalfUNetSettings:
        return self._settings

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        e1 = self.enc1(x)
        e2 = self.enc2(self.pool(e1))
        e3 = self.enc3(self.pool(e2))
        s = e1 + self.up2(e2)[..., : e1.shape[-2], : e1.shape[-1]] + self.up4(e3)[..., : e1.shape[-2], : e1.shape[-1]]
        return self.head(self.dec(s))
