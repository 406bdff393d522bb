"""Run-time configuration of the StyleGAN2 graph (reference constants.py:1-25, made configurable).

The reference hard-codes 256 px / batch 4; here every value can be overridden through the
environment (``L2I_RESOLUTION``, ``L2I_BATCH_SIZE``, ``L2I_DTYPE``, ``L2I_G_PATH``, ``L2I_REG_PATH``)
or by assigning to this module before the graph is built.
"""
import os

BATCH_SIZE = int(os.environ.get("L2I_BATCH_SIZE", 4))
DIM_Z = 512
resolution = int(os.environ.get("L2I_RESOLUTION", 256))
useGPU = True
NUM_CHANNELS = 3

# checkpoints: rosinality-format generator ({'g_ema': state_dict}) and ResNet-50 regressor ({'model': ...})
reg_json = None
reg_path = os.environ.get("L2I_REG_PATH", "/path/003_dict.model")
g_path = os.environ.get("L2I_G_PATH", "/path/550000.pt")
# torchvision vgg19 state_dict for the content loss (the reference downloads it: models.vgg19(pretrained=True))
vgg_path = os.environ.get("L2I_VGG_PATH", "")
# A missing checkpoint is an error, as in the reference (torch.load raises).  Benchmarks and tests, which run on random-init
# weights of the real architecture, opt in explicitly: L2I_ALLOW_RANDOM_INIT=1 or --allow_random_init.
allow_random_init = os.environ.get("L2I_ALLOW_RANDOM_INIT", "0") not in ("0", "")
compute_dtype = "fp32" if os.environ.get("L2I_DTYPE", "bf16").lower() in ("fp32", "f32", "float32") else "bf16"
walk_is_mlp = False
# stock ResNet-50 regressor under bf16 autocast + channels_last (train.py --amp); False = the reference's fp32 arithmetic
# fold the frozen eval-mode BatchNorms of the regressor into its convolutions (exact up to fp32 rounding; L2I_REG_FOLD_BN=0 disables)
reg_fold_bn = os.environ.get("L2I_REG_FOLD_BN", "1") not in ("0", "")
reg_amp = os.environ.get("L2I_REG_AMP", "0") not in ("0", "")
