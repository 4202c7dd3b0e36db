"""gmlm_b200 — B200-native (sm_100a) message passing for the GNN encoder of chungimungi/GMLM.

Drop-in for the three torch_geometric names the reference imports (``main.py:6-7``)::

    from gmlm_b200 import RGCNConv, GraphNorm      # instead of torch_geometric.nn
    from gmlm_b200 import degree                   # instead of torch_geometric.utils

plus the fused forms of the reference's own helpers on the path
(``soft_masking_gnn_input`` main.py:92-99, ``edge_type_from_degree`` main.py:253-267,
``GraphEncoder.get_graph_embeddings`` main.py:250-320).

All arithmetic runs in ``libgmlm_b200.so`` (hand-written CUDA behind the C ABI of
``include/gmlm_b200.h``).  There is no CPU fallback: CPU tensors raise, and a missing
library raises with build instructions.
"""
from ._lib import GmlmError, lib_path, load as load_library, set_tuning  # noqa: F401
from .graph import CSR, RelGraph, build_csr, clear_graph_cache, get_rel_graph, transpose_csr  # noqa: F401
from .ops import (  # noqa: F401
    degree,
    edge_type_from_degree,
    graph_norm,
    layer_norm,
    rgcn_aggregate,
    soft_masking_gnn_input,
    spmm,
)
from .nn import GraphNorm, RGCNConv  # noqa: F401
from .encoder import GraphEncoder, MultiScaleFusion  # noqa: F401
from .graphed import GraphedEncoderStep  # noqa: F401
from .losses import nt_xent_loss  # noqa: F401
from .sampling import generate_active_node_mask, weighted_sample_without_replacement  # noqa: F401
from .dist_norm import partitioned_graph_norm  # noqa: F401
from .dist_encoder import CudaPartitionOps, PartitionedGraphEncoder, sync_gradients  # noqa: F401
from .attn import GATConv, GCNConv, LoopGraph, gat_aggregate, gat_dropout_mask, get_loop_graph  # noqa: F401
from .ingest import (augment_graph, build_dropped_graph, edge_dropout_mask, load_npz_graph, load_rel_graph,  # noqa: F401
                     save_rel_graph)

__version__ = "0.1.0"
