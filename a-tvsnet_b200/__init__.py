"""a-tvsnet_b200 - B200 (sm_100a) implementation of the A-TVSNet inference hot path behind
the reference's own Python operator surface.  Import as ``atvsnet_b200`` (root-level shim:
the directory name contains a hyphen).  Everything runs on CUDA tensors through
libatvs.so; there is no CPU fallback."""
from . import _lib, variables  # noqa: F401
from .flags import FLAGS  # noqa: F401
from .homography_warping import get_homographies, homography_warping, homography_warping_by_depth  # noqa: F401
from .model import (TVSNet, TVSNet_base, TVSNet_base_siamese, TVSNet_feature_extraction, TVSNet_refine,  # noqa: F401
                    build_cost_volume, cost_volume_aggregation, cost_volume_aggregation_refine, cost_volume_reasoning,
                    output_conv, output_conv_refine, prob2depth, prob2depth_upsample)
from .atvsnet import (AttAggregation, AttAggregation_keepchannel, AttAggregation_refine,  # noqa: F401
                      AttAggregation_refine_keepchannel, OutputConv, OutputConv_refine, StackedUNet,
                      StackedUNet_prob)
from .network import Network  # noqa: F401
from . import ckpt, eval_errors, fem, fusion, pipeline, preprocess, refine, synthetic  # noqa: F401
