"""vlsat_b200 - B200-native implementation of the VL-SAT (wz7in/CVPR2023-VLSAT) hot path.

Python/PyTorch host modules with the reference's class names, constructor arguments, ``state_dict``
layout and forward signatures; all arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI
in ``include/vlsat_b200.h`` (``libvlsat_b200.so``). No CPU / PyTorch fallback.
"""
from .attention import MultiHeadAttention, ScaledDotProductAttention, SceneContext          # noqa: F401
from .gat import (Aggre_Index, Gen_Index, GraphContext, GraphEdgeAttenNetwork, MLP,          # noqa: F401
                  MultiHeadedEdgeAttention, build_mlp)
from .mmg import MMG, GraphEdgeAttenNetworkLayers                                            # noqa: F401
from .mmgnet import (AdapterModel, DEFAULT_MODEL_CONFIG, Mmgnet, accelerate_reference_model,  # noqa: F401
                     adopt_parameters, load_model_config)
from .pointnet import PointNetfeat, PointNetRelClsMulti                                      # noqa: F401
from .graph import GraphedForward, GraphedTrainStep, StreamedInference                                                           # noqa: F401
from . import autograd, data_prep, eval_ranks, ops, synth, train_glue, train_path                                   # noqa: F401

__version__ = "0.1.0"
