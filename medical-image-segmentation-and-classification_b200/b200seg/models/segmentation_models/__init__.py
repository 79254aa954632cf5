from .AttentionUNet import AttentionUNet
from .R2U_Net import R2U_Net
from .R2AttU_Net import R2AttU_Net
from .ResnetUnet import ResNetUnet
