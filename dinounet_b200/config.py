"""Static configuration of the Dino U-Net variants (host-side mirror of the reference registries).

Reference: DINOv3_MODEL_FACTORIES / _INTERACTION_INDEXES / _MODEL_INFO (dinounet_training.py:29-48) and the hub
factories dinounet/dinov3/hub/backbones.py:201-237 (vits16), 279-315 (vitb16), 318-371 (vitl16), 452-494 (vit7b16).
"""
from dataclasses import dataclass
from typing import Dict, Tuple


@dataclass(frozen=True)
class VariantConfig:
    name: str
    embed_dim: int
    depth: int
    num_heads: int
    ffn_layer: str           # "mlp" | "swiglu64"
    ffn_hidden: int
    qkv_bias: bool
    interaction_indexes: Tuple[int, ...]
    untie_global_and_local_cls_norm: bool = False
    drop_path_rate: float = 0.0
    params: str = ""


VARIANTS: Dict[str, VariantConfig] = {
    "dinounet_s": VariantConfig("dinounet_s", 384, 12, 6, "mlp", 1536, True, (2, 5, 8, 11), params="~22M"),
    "dinounet_b": VariantConfig("dinounet_b", 768, 12, 12, "mlp", 3072, True, (2, 5, 8, 11), params="~86M"),
    "dinounet_l": VariantConfig("dinounet_l", 1024, 24, 16, "mlp", 4096, True, (4, 11, 17, 23), params="~300M"),
    "dinounet_7b": VariantConfig("dinounet_7b", 4096, 40, 32, "swiglu64", 8192, False, (9, 19, 29, 39), True, 0.4,
                                 params="~7B"),
}

# checkpoint names the reference trainers expect (dinounet_training.py:893,905,917,930)
CHECKPOINTS = {
    "dinounet_s": "dinounet/checkpoints/dinov3_vits16_pretrain_lvd1689m-08c60483.pth",
    "dinounet_b": "dinounet/checkpoints/dinov3_vitb16_pretrain_lvd1689m-73cec8be.pth",
    "dinounet_l": "dinounet/checkpoints/dinov3_vitl16_pretrain_lvd1689m-8aa4cbdd.pth",
    "dinounet_7b": "dinounet/checkpoints/dinov3_vit7b16_pretrain_lvd1689m-a955f4ea.pth",
}

PATCH_SIZE = 16
N_STORAGE_TOKENS = 4
N_PREFIX = 1 + N_STORAGE_TOKENS
ROPE_BASE = 100.0
LN_EPS_VIT = 1e-5        # "layernormbf16" (vision_transformer.py:29)
LN_EPS_ADAPTER = 1e-6    # partial(nn.LayerNorm, eps=1e-6) (dinov3_adapter.py:368)
BN_EPS = 1e-5
IN_EPS = 1e-5

# adapter hyper-parameters fixed by the reference (dinounet_training.py:754-769)
CONV_INPLANE = 64
DEFORM_HEADS = 16
DEFORM_POINTS = 4
DEFORM_RATIO = 0.5
CFFN_RATIO = 0.25
FAPM_RANK = 256          # dinounet_training.py:449

# what the nnU-Net planner emits for main_dinov3's forced 2d / 512x512 / 4-stage plan
DEFAULT_ARCHITECTURE = {
    "n_stages": 4,
    "features_per_stage": [32, 64, 128, 256],
    "conv_op": "torch.nn.modules.conv.Conv2d",
    "kernel_sizes": [[3, 3]] * 4,
    "strides": [[1, 1], [2, 2], [2, 2], [2, 2]],
    "n_conv_per_stage": [2, 2, 2, 2],
    "n_conv_per_stage_decoder": [2, 2, 2],
    "conv_bias": True,
    "norm_op": "torch.nn.modules.instancenorm.InstanceNorm2d",
    "norm_op_kwargs": {"eps": 1e-5, "affine": True},
    "dropout_op": None,
    "dropout_op_kwargs": None,
    "nonlin": "torch.nn.LeakyReLU",
    "nonlin_kwargs": {"inplace": True},
}
